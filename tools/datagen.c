/*
 * Deterministic synthetic FASTQ-like / SAM-like text, byte-for-byte the generator that SURVEY.md
 * Appendix B declares normative for every number quoted on this path (splitmix64, 4 Mi-base genome,
 * 150-base reads, 2-state Markov qualities).  Used by tests/, bench.py and the tools; it is a workload
 * generator, not part of the codec.
 *
 *   library:  size_t b200gen_fill(int kind, uint64_t seed, uint8_t *dst, size_t nbytes)
 *             kind 0 = FASTQ-like, 1 = SAM-like, 2 = BAM-like; writes exactly nbytes (the stream is cut mid-record).
 *   cli:      datagen fastq|sam|bam <bytes> <seed>   (stops after the first record reaching <bytes>,
 *             like the survey probe, so the md5 sums in SURVEY.md Appendix B reproduce)
 * BAM-like (BASELINE config 5: what htslib hands to bgzf_compress when samtools writes BAM): the SAM-like records in
 * BAM's binary layout — u32 block size, 32-byte little-endian core, NUL-terminated name, one CIGAR word, 4-bit packed
 * bases, raw Phred bytes, NM/AS/RG tags.  Not a valid BAM file (no header, no bin/index fields of any meaning); it has
 * BAM's byte statistics.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GENOME_LEN (4u << 20)
#define READ_LEN 150

typedef struct {
    uint64_t s;
    char *genome;
} gen_t;

static inline uint64_t draw(gen_t *g)
{
    uint64_t z = (g->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static const char BASES[4] = { 'A', 'C', 'G', 'T' };

static void make_genome(gen_t *g)
{
    g->genome = (char *)malloc(GENOME_LEN);
    for (uint32_t i = 0; i < GENOME_LEN; i += 32) {
        uint64_t r = draw(g);
        for (int k = 0; k < 32; k++)
            g->genome[i + k] = BASES[(r >> (2 * k)) & 3];
    }
}

static void make_read(gen_t *g, char *seq, char *qual, uint32_t pos)
{
    static const char QSYM[4] = { 'F', ':', ',', '#' };
    for (int k = 0; k < READ_LEN; k++) {
        char c = g->genome[(pos + k) % GENOME_LEN];
        uint64_t r = draw(g);
        if ((r & 1023) < 5)
            c = BASES[(r >> 10) & 3];
        if (((r >> 20) & 4095) == 0)
            c = 'N';
        seq[k] = c;
    }
    int st = 0;
    for (int k = 0; k < READ_LEN; k++) {
        uint32_t r = (uint32_t)(draw(g) & 0xffff);
        if (st == 0)
            st = r < 0xEE00 ? 0 : r < 0xF800 ? 1 : r < 0xFE00 ? 2 : 3;
        else
            st = r < 0x8000 ? 0 : r < 0xC000 ? 1 : r < 0xE800 ? 2 : 3;
        if (k > READ_LEN - 20 && (r & 7) == 0)
            st = 2;
        qual[k] = QSYM[st];
    }
}

/* Emits records through sink() until it returns non-zero. */
typedef int (*sink_fn)(void *ctx, const char *rec, size_t len);

static void put32(unsigned char *p, uint32_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); p[2] = (unsigned char)(v >> 16); p[3] = (unsigned char)(v >> 24); }
static void put16(unsigned char *p, uint32_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); }

static int bam_record(char *rec, const char *name, int flag, uint32_t pos, int mapq, uint32_t pnext, int tlen, const char *seq, const char *qual,
                      uint32_t nm, uint32_t as)
{
    unsigned char *o = (unsigned char *)rec + 4;
    const uint32_t lname = (uint32_t)strlen(name) + 1;
    put32(o, 0); put32(o + 4, pos); o[8] = (unsigned char)lname; o[9] = (unsigned char)mapq; put16(o + 10, 4681); put16(o + 12, 1);
    put16(o + 14, (uint32_t)flag); put32(o + 16, READ_LEN); put32(o + 20, 0); put32(o + 24, pnext); put32(o + 28, (uint32_t)tlen);
    o += 32;
    memcpy(o, name, lname); o += lname;
    put32(o, (uint32_t)READ_LEN << 4); o += 4;
    for (int i = 0; i < READ_LEN; i += 2) {
        int a = seq[i] == 'A' ? 1 : seq[i] == 'C' ? 2 : seq[i] == 'G' ? 4 : seq[i] == 'T' ? 8 : 15;
        int b = i + 1 < READ_LEN ? (seq[i + 1] == 'A' ? 1 : seq[i + 1] == 'C' ? 2 : seq[i + 1] == 'G' ? 4 : seq[i + 1] == 'T' ? 8 : 15) : 0;
        *o++ = (unsigned char)((a << 4) | b);
    }
    for (int i = 0; i < READ_LEN; i++) *o++ = (unsigned char)(qual[i] - 33);
    memcpy(o, "NMC", 3); o[3] = (unsigned char)nm; o += 4;
    memcpy(o, "ASC", 3); o[3] = (unsigned char)as; o += 4;
    memcpy(o, "RGZgrp1", 8); o += 8;
    const uint32_t n = (uint32_t)(o - (unsigned char *)rec);
    put32((unsigned char *)rec, n - 4);
    return (int)n;
}

static void generate(int sam, uint64_t seed, sink_fn sink, void *ctx)
{
    gen_t g;
    g.s = seed;
    make_genome(&g);
    char seq[READ_LEN + 8], qual[READ_LEN + 8], rec[1024];
    uint64_t i = 0;
    uint32_t pos = 0;
    int stop = 0;
    if (sam == 1) {
        int n = snprintf(rec, sizeof rec,
                         "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chrSim\tLN:%u\n@PG\tID:b200sim\tPN:b200sim\n",
                         GENOME_LEN);
        stop = sink(ctx, rec, (size_t)n);
    }
    while (!stop) {
        uint32_t tile = 1101 + (uint32_t)((i / 20000) % 78), lane = 1 + (uint32_t)((i / 1560000) % 8);
        uint32_t x = 1000 + (uint32_t)(draw(&g) % 30000);
        uint32_t y = 1000 + (uint32_t)((i % 20000) * 10 + (draw(&g) % 10));
        int n;
        if (!sam) {
            pos = (uint32_t)(draw(&g) % GENOME_LEN);
            make_read(&g, seq, qual, pos);
            n = snprintf(rec, sizeof rec, "@B200SIM:7:HXXCCXY22:%u:%u:%u:%u 1:N:0:ACGTACGT\n%.*s\n+\n%.*s\n",
                         lane, tile, x, y, READ_LEN, seq, READ_LEN, qual);
        } else {
            pos += (uint32_t)(draw(&g) % 10);
            if (pos >= GENOME_LEN - READ_LEN)
                pos = 0;
            make_read(&g, seq, qual, pos);
            uint32_t r = (uint32_t)draw(&g);
            int flag = (r & 1) ? 99 : 147;
            int mapq = (r & 0xf0) ? 60 : (int)((r >> 8) % 60);
            int tlen = 300 + (int)((r >> 16) % 200);
            if (flag == 147)
                tlen = -tlen;
            uint32_t pnext = flag == 99 ? pos + tlen - READ_LEN : pos + tlen + READ_LEN;
            if (sam == 2) {
                char name[96];
                snprintf(name, sizeof name, "B200SIM:7:HXXCCXY22:%u:%u:%u:%u", lane, tile, x, y);
                n = bam_record(rec, name, flag, pos, mapq, pnext, tlen, seq, qual, (r >> 24) & 3, READ_LEN - ((r >> 24) & 3) * 5);
            } else
            n = snprintf(rec, sizeof rec,
                         "B200SIM:7:HXXCCXY22:%u:%u:%u:%u\t%d\tchrSim\t%u\t%d\t%dM\t=\t%u\t%d\t%.*s\t%.*s\t"
                         "NM:i:%u\tMD:Z:%d\tAS:i:%u\tXS:i:%u\tRG:Z:grp%u\n",
                         lane, tile, x, y, flag, pos + 1, mapq, READ_LEN, pnext + 1, tlen, READ_LEN, seq, READ_LEN,
                         qual, (r >> 24) & 3, READ_LEN, READ_LEN - ((r >> 24) & 3) * 5, (r >> 26) * 2, lane);
        }
        stop = sink(ctx, rec, (size_t)n);
        i++;
    }
    free(g.genome);
}

typedef struct {
    uint8_t *dst;
    size_t cap, n;
} fill_ctx;

static int fill_sink(void *c, const char *rec, size_t len)
{
    fill_ctx *f = (fill_ctx *)c;
    size_t room = f->cap - f->n;
    size_t take = len < room ? len : room;
    memcpy(f->dst + f->n, rec, take);
    f->n += take;
    return f->n >= f->cap;
}

size_t b200gen_fill(int kind, uint64_t seed, uint8_t *dst, size_t nbytes)
{
    fill_ctx f = { dst, nbytes, 0 };
    if (nbytes)
        generate(kind, seed, fill_sink, &f);
    return f.n;
}

#ifdef DATAGEN_MAIN
typedef struct {
    uint64_t target, n;
} cli_ctx;

static int cli_sink(void *c, const char *rec, size_t len)
{
    cli_ctx *f = (cli_ctx *)c;
    fwrite(rec, 1, len, stdout);
    f->n += len;
    return f->n >= f->target;
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s fastq|sam|bam <bytes> <seed>\n", argv[0]);
        return 2;
    }
    cli_ctx c = { strtoull(argv[2], 0, 10), 0 };
    setvbuf(stdout, 0, _IOFBF, 1 << 22);
    generate(!strcmp(argv[1], "sam") ? 1 : !strcmp(argv[1], "bam") ? 2 : 0, strtoull(argv[3], 0, 10), cli_sink, &c);
    return 0;
}
#endif
