/*
 * applet_7bgzf.c — the `7bgzf` command line of the reference (applet/7bgzf.c:371-551), driving the GPU codec.
 *
 *   7bgzf -c -l6 [-@ N] < in > out.bgz        compress (stdin -> stdout only)
 *   7bgzf -d [-@ N] < in.bgz > out            decompress
 *
 * Same flags as the reference's popt table (7bgzf.c:379-398): -l[N]/--libdeflate[=N] (1-12, default 6),
 * -z -m -s -S -n -C -i -K -T [N] and -Z N (the other method flags: each selects a level for the one GPU codec,
 * there is no multi-backend dispatch), -@ N/--threads (accepted; the GPU path is always block-parallel),
 * -d/--decompress, -c/--stdout (ignored).  Exactly one method unless -d (7bgzf.c:467-479); refuses a
 * terminal on stdin or stdout (:491,498); stderr carries the reference's lines:
 *   "compression level = %d (libdeflate)", "%d done.", "ellapsed time: %.6f sec".
 * Payload blocks are always 0xff00 bytes (the reference's multi-thread rule, 7bgzf.c:141-147).
 * Host code is C; CUDA is reached only through the b200bgzf_* C ABI.  No GPU => error, no CPU fallback.
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include "../../include/b200bgzf.h"

#define CHUNK_BLOCKS 4096u /* 255 MiB of payload per API call */

struct method_flag {
    char short_opt;
    const char *long_opt;
    const char *label;
    int default_level;
    int max_level;
};

static const struct method_flag k_flags[] = {
    { 'z', "zlib", "zlib", 6, 9 },       { 'm', "miniz", "miniz", 1, 9 },      { 's', "slz", "slz", 1, 1 },
    { 'l', "libdeflate", "libdeflate", 6, 12 }, { 'S', "7zip", "7zip", 2, 9 }, { 'n', "zlibng", "zlibng", 6, 9 },
    { 'C', "cryptopp", "cryptopp", 6, 9 }, { 'i', "igzip", "igzip", 1, 4 },    { 'K', "kzip", "kzip", 1, 1 },
    { 'Z', "zopfli", "zopfli", 0, 12 },   { 'T', "store", "store", 1, 1 },
};
#define NFLAGS (sizeof k_flags / sizeof k_flags[0])

static void usage(const char *argv0)
{
    fprintf(stderr,
            "Usage: %s -cz9 < dec.bin > enc.bgz or -cd < enc.bgz > dec.bin\n"
            "  -c, --stdout              stdout (currently ignored; always output to stdout)\n"
            "  -z, --zlib[=level]        1-9 (default 6) zlib\n"
            "  -m, --miniz[=level]       1-9 (default 1) miniz\n"
            "  -s, --slz[=level]         1-1 (default 1) slz\n"
            "  -l, --libdeflate[=level]  1-12 (default 6) libdeflate\n"
            "  -S, --7zip[=level]        1-9 (default 2) 7zip\n"
            "  -n, --zlibng[=level]      1-9 (default 6) zlibng\n"
            "  -C, --cryptopp[=level]    1-9 (default 6) cryptopp\n"
            "  -i, --igzip[=level]       1-4 (default 1) igzip\n"
            "  -K, --kzip[=level]        1-1 (default 1) kzip\n"
            "  -Z, --zopfli=numiterations zopfli\n"
            "  -T, --store[=level]       1-1 (default 1) store\n"
            "  -@, --threads=threads     threads\n"
            "  -d, --decompress          decompress\n"
            "\nNote: every method runs on the B200 BGZF codec (libdeflate level classes 1-12).\n",
            argv0);
}

static int read_full(FILE *f, unsigned char *buf, size_t want, size_t *got)
{
    size_t n = 0;
    while (n < want) {
        size_t r = fread(buf + n, 1, want - n, f);
        if (r == 0) break;
        n += r;
    }
    *got = n;
    return ferror(f) ? -1 : 0;
}

static int do_compress(b200bgzf_ctx *ctx, FILE *in, FILE *out, int level)
{
    const size_t chunk = (size_t)CHUNK_BLOCKS * B200BGZF_BLOCK_SIZE;
    const size_t bound = b200bgzf_compress_bound(chunk, B200BGZF_BLOCK_SIZE);
    unsigned char *ibuf = (unsigned char *)malloc(chunk), *obuf = (unsigned char *)malloc(bound);
    if (!ibuf || !obuf) { fprintf(stderr, "out of memory\n"); return 1; }
    long blocks = 0;
    int ret = 0;
    for (;;) {
        size_t got = 0, produced = 0;
        if (read_full(in, ibuf, chunk, &got) != 0) { ret = 1; break; }
        if (got == 0) break;
        int r = b200bgzf_compress_host(ctx, ibuf, got, B200BGZF_BLOCK_SIZE, level, obuf, bound, &produced, 0);
        if (r != 0) {
            if (r == B200BGZF_E_NOFIT) fprintf(stderr, "libdeflate_deflate %d\n", 1);
            else fprintf(stderr, "b200bgzf: %s (%s)\n", b200bgzf_strerror(r), b200bgzf_last_error(ctx));
            ret = 1;
            break;
        }
        if (fwrite(obuf, 1, produced, out) != produced) { ret = 1; break; }
        blocks += (long)((got + B200BGZF_BLOCK_SIZE - 1) / B200BGZF_BLOCK_SIZE);
        fprintf(stderr, "%ld\r", blocks);
        if (got < chunk) break;
    }
    if (!ret) {
        /* EOF marker (7bgzf.c:283-289) */
        static const unsigned char eof[28] = { 0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                               0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        fwrite(eof, 1, 28, out);
        fprintf(stderr, "%ld done.\n", blocks);
    }
    free(ibuf);
    free(obuf);
    return ret;
}

static int do_decompress(b200bgzf_ctx *ctx, FILE *in, FILE *out)
{
    /* read members until ~256 MiB of compressed data is buffered, inflate, repeat */
    const size_t cap = 256u << 20;
    unsigned char *ibuf = (unsigned char *)malloc(cap + B200BGZF_MAX_BLOCK_SIZE);
    size_t ocap = 0;
    unsigned char *obuf = NULL;
    if (!ibuf) { fprintf(stderr, "out of memory\n"); return 1; }
    size_t have = 0;
    long members = 0;
    int ret = 0, eof_seen = 0;
    while (!ret) {
        if (!eof_seen) {
            size_t got = 0;
            if (read_full(in, ibuf + have, cap - have, &got) != 0) { ret = 1; break; }
            have += got;
            if (have < cap) eof_seen = 1;
        }
        if (have == 0) break;
        /* whole members only */
        size_t used = 0, total = 0, nm = 0;
        while (used + 18 <= have) {
            const unsigned char *p = ibuf + used;
            if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4) || p[10] != 6 || p[12] != 'B' || p[13] != 'C') {
                fprintf(stderr, "not BGZF or corrupted\n");
                ret = -1;
                break;
            }
            size_t sz = (size_t)(p[16] | (p[17] << 8)) + 1;
            if (used + sz > have) break;
            const unsigned char *t = p + sz - 4;
            total += (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
            used += sz;
            nm++;
        }
        if (ret) break;
        if (used == 0) {
            if (eof_seen) { fprintf(stderr, "not BGZF or corrupted\n"); ret = -1; }
            break;
        }
        if (total > ocap) {
            free(obuf);
            ocap = total + (total >> 2) + 65536;
            obuf = (unsigned char *)malloc(ocap);
            if (!obuf) { fprintf(stderr, "out of memory\n"); ret = 1; break; }
        }
        size_t produced = 0;
        int r = b200bgzf_inflate_host(ctx, ibuf, used, obuf, ocap, &produced, 0);
        if (r != 0) {
            fprintf(stderr, "inflate %d\n", r);
            ret = 1;
            break;
        }
        if (fwrite(obuf, 1, produced, out) != produced) { ret = 1; break; }
        members += (long)nm;
        fprintf(stderr, "%ld\r", members);
        memmove(ibuf, ibuf + used, have - used);
        have -= used;
        if (eof_seen && have == 0) break;
    }
    if (!ret) fprintf(stderr, "%ld done.\n", members);
    free(ibuf);
    free(obuf);
    return ret;
}

int main(int argc, char **argv)
{
    int levels[NFLAGS];
    int decompress = 0, nthreads = 1, bad = 0;
    memset(levels, 0, sizeof levels);
    /* allow `cielbox 7bgzf ...` style invocation */
    if (argc > 1 && !strcmp(argv[1], "7bgzf")) { argv++; argc--; }

    static const struct option longopts[] = {
        { "stdout", no_argument, 0, 'c' },       { "zlib", optional_argument, 0, 'z' },
        { "miniz", optional_argument, 0, 'm' },  { "slz", optional_argument, 0, 's' },
        { "libdeflate", optional_argument, 0, 'l' }, { "7zip", optional_argument, 0, 'S' },
        { "zlibng", optional_argument, 0, 'n' }, { "cryptopp", optional_argument, 0, 'C' },
        { "igzip", optional_argument, 0, 'i' },  { "kzip", optional_argument, 0, 'K' },
        { "zopfli", required_argument, 0, 'Z' }, { "store", optional_argument, 0, 'T' },
        { "threads", required_argument, 0, '@' }, { "decompress", no_argument, 0, 'd' },
        { "help", no_argument, 0, 'h' },         { 0, 0, 0, 0 },
    };
    int opt;
    while ((opt = getopt_long(argc, argv, "cz::m::s::l::S::n::C::i::K::Z:T::@:dh", longopts, NULL)) != -1) {
        if (opt == 'c') continue;
        if (opt == 'd') { decompress = 1; continue; }
        if (opt == '@') { nthreads = atoi(optarg); continue; }
        if (opt == 'h' || opt == '?') { bad = 1; continue; }
        for (size_t k = 0; k < NFLAGS; k++)
            if (k_flags[k].short_opt == opt)
                levels[k] = optarg ? (int)strtol(optarg, NULL, 10) : k_flags[k].default_level;
    }
    (void)nthreads;
    int chosen = -1, nchosen = 0, level_sum = 0;
    for (size_t k = 0; k < NFLAGS; k++)
        if (levels[k]) { chosen = (int)k; nchosen++; level_sum += levels[k]; }
    if (bad || (!decompress && nchosen != 1) || (decompress && nchosen != 0)) {
        usage(argv[0]);
        return 1;
    }
    if (isatty(fileno(stdin)) || isatty(fileno(stdout))) {
        usage(argv[0]);
        return -1;
    }
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    b200bgzf_ctx *ctx = NULL;
    const char *dev = getenv("B200BGZF_DEVICE");
    int r = b200bgzf_create(&ctx, dev && *dev ? atoi(dev) : -1);
    if (r != 0) {
        fprintf(stderr, "b200bgzf: cannot initialise the GPU codec: %s\n", b200bgzf_strerror(r));
        return 1;
    }
    int ret;
    if (decompress) {
        ret = do_decompress(ctx, stdin, stdout);
    } else {
        int level = level_sum;
        fprintf(stderr, "compression level = %d (%s)\n", level_sum, k_flags[chosen].label);
        if (level < 1) level = 1;
        if (level > 12) level = 12;
        ret = do_compress(ctx, stdin, stdout, level);
    }
    fflush(stdout);
    b200bgzf_destroy(ctx);
    gettimeofday(&t1, NULL);
    fprintf(stderr, "ellapsed time: %.6f sec\n", (t1.tv_sec + t1.tv_usec * 0.000001) - (t0.tv_sec + t0.tv_usec * 0.000001));
    return ret;
}
