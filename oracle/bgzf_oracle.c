/*
 * bgzf_oracle.c — CPU restatement of the reference's BGZF hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * (as liboracle.so).  The product (7bgzf_b200/) never links, calls or falls back to anything in oracle/.
 *
 * What is restated here, plainly and sequentially, each from the reference lines cited:
 *   oracle_crc32            CRC-32 of the payload                    lib/zlib/crc32.c:1015 (hook), libdeflate_crc32.c:255-262
 *   oracle_crc32_combine    crc(A||B) from crc(A), crc(B), |B|       lib/zlib/crc32.c:155-186,1021-1049
 *   oracle_eof_block        the 28-byte EOF member                   bgzf_compress.c:43-49, applet/7bgzf.c:283-289
 *   oracle_bgzf_frame       header + BSIZE + payload + CRC + ISIZE   bgzf_compress.c:191-196, applet/7bgzf.c:263-272
 *   oracle_store_deflate    stored-block "compressor"                lib/zlibutil.c:302-325
 *   oracle_read_gz_header   member header parser (BGZF/MiGz/mgzip)   applet/7bgzf.c:81-131
 *   oracle_inflate          raw DEFLATE decoder                      lib/libdeflate/deflate_decompress.c:721-1004 (tables),
 *                                                                   decompress_template.h:44-772 (blocks, decode loop)
 *   oracle_bgzf_decompress  the applet's decompress loop             applet/7bgzf.c:295-365
 *   oracle_parse_method     BGZF_METHOD=<name><digits>               bgzf_compress.c:53-113
 *   oracle_passthrough      tiny-input rule of the encoder           lib/libdeflate/deflate_compress.c:3919,4035
 *
 * The compressor itself (libdeflate's lazy/near-optimal parsers) is NOT restated: its output is not what
 * parity is defined on (sizes may differ by <= 3 %); the compiled reference in oracle/_ref is the size
 * baseline and the decoder of record.  Parity pins: every function above is checked in tests/ against
 * oracle/_ref (the unmodified reference built by oracle/Makefile.ref) and against the survey's known answers.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

/* ---------------------------------------------------------------- CRC-32 ---------------------------------- */

static uint32_t crc_table[256];
static int crc_ready;

static void crc_init(void)
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t r = i;
        for (int k = 0; k < 8; k++) r = (r & 1) ? (r >> 1) ^ 0xEDB88320u : r >> 1;
        crc_table[i] = r;
    }
    crc_ready = 1;
}

uint32_t oracle_crc32(uint32_t crc, const uint8_t *buf, size_t len)
{
    if (!crc_ready) crc_init();
    crc = ~crc;
    for (size_t i = 0; i < len; i++) crc = crc_table[(crc ^ buf[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

/* a(x)*b(x) mod p(x), reflected: the reference's multmodp (crc32.c:155-171) */
static uint32_t multmodp(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = b & 1 ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}

/* x^(8*len2) mod p(x) by square-and-multiply (the job of x2nmodp, crc32.c:177-186) */
static uint32_t xpow8(uint64_t len2)
{
    uint32_t p = 1u << 31, sq = multmodp(1u << 30, 1u << 30); /* x^2 */
    sq = multmodp(sq, sq);                                    /* x^4 */
    sq = multmodp(sq, sq);                                    /* x^8 */
    while (len2) {
        if (len2 & 1) p = multmodp(sq, p);
        sq = multmodp(sq, sq);
        len2 >>= 1;
    }
    return p;
}

uint32_t oracle_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
    return multmodp(xpow8(len2), crc1) ^ crc2;
}

/* ---------------------------------------------------------------- framing --------------------------------- */

static void put16(uint8_t *p, uint32_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; }
static void put32(uint8_t *p, uint32_t v) { put16(p, v); put16(p + 2, v >> 16); }
static uint32_t get16(const uint8_t *p) { return p[0] | (p[1] << 8); }
static uint32_t get32(const uint8_t *p) { return get16(p) | ((uint32_t)get16(p + 2) << 16); }

static const uint8_t k_hdr16[16] = { 0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 'B', 'C', 0x02, 0 };

size_t oracle_eof_block(uint8_t *dst)
{
    memcpy(dst, k_hdr16, 16);
    put16(dst + 16, 27);
    dst[18] = 0x03;
    dst[19] = 0;
    memset(dst + 20, 0, 8);
    return 28;
}

/* wraps an already-deflated payload; returns the member size (18 + dlen + 8) */
size_t oracle_bgzf_frame(uint8_t *dst, const uint8_t *deflated, size_t dlen, const uint8_t *src, size_t slen)
{
    memcpy(dst, k_hdr16, 16);
    put16(dst + 16, (uint32_t)(dlen + 25));
    memmove(dst + 18, deflated, dlen);
    put32(dst + 18 + dlen, oracle_crc32(0, src, slen));
    put32(dst + 22 + dlen, (uint32_t)slen);
    return dlen + 26;
}

int oracle_store_deflate(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t sourceLen)
{
    size_t blocks = (sourceLen + 65534) / 65535;
    if (*destLen < sourceLen + 5 * blocks) return -5; /* Z_BUF_ERROR */
    *destLen = 0;
    for (size_t i = 0; i < blocks; i++) {
        uint32_t n = sourceLen < 65535 ? (uint32_t)sourceLen : 65535u;
        dest[0] = i < blocks - 1 ? 0 : 1;
        put16(dest + 1, n);
        put16(dest + 3, ~n);
        memcpy(dest + 5, source, n);
        source += n;
        sourceLen -= n;
        dest += n + 5;
        *destLen += n + 5;
    }
    return 0;
}

/* encoder's tiny-input rule: inputs of at most 55-4*level bytes are stored */
int oracle_passthrough(int level) { return 55 - 4 * level; }

/* returns the header length (offset of the DEFLATE data) or 0; *block_len = whole member size */
int oracle_read_gz_header(const uint8_t *data, int size, int *extra_off, int *extra_len, int *block_len)
{
    int n, flags;
    if (size < 4 || data[0] != 0x1f || data[1] != 0x8b) return 0;
    flags = data[3];
    if (data[2] != 8 || (flags & 0xE0)) return 0;
    n = 10;
    *extra_off = n + 2;
    *extra_len = 0;
    *block_len = 0;
    if (flags & 0x04) {
        if (size < n + 2) return 0;
        int len = (int)get16(data + n);
        n += 2;
        *extra_off = n;
        *extra_len = len;
        if (size < n + len) return 0;
        n += len;
    }
    if (flags & 0x08) while (n < size && data[n++]) {}
    if (flags & 0x10) while (n < size && data[n++]) {}
    if (flags & 0x02) {
        if (n + 2 > size) return 0;
        n += 2;
    }
    const uint8_t *x = data + *extra_off;
    if (*extra_len == 6 && !memcmp(x, "BC\x02\x00", 4)) *block_len = (int)get16(x + 4) + 1;
    else if (*extra_len == 8 && !memcmp(x, "MZ\x04\x00", 4)) *block_len = (int)get32(x + 4) + n + 8;
    else if (*extra_len == 20 && !memcmp(x, "IG\x10\x00", 4)) *block_len = (int)get32(x + 4);
    else if (*extra_len == 8 && !memcmp(x, "IG\x04\x00", 4)) *block_len = (int)get32(x + 4);
    else if (*extra_len == 4 && x[3] == 0x7d) *block_len = (int)(get32(x) & 0xffffff);
    else return 0;
    return n;
}

/* ---------------------------------------------------------------- inflate --------------------------------- */

#define OR_OK 0
#define OR_BAD_DATA 1
#define OR_SHORT_OUTPUT 2
#define OR_INSUFFICIENT_SPACE 3

#define LITLEN_BITS 11
#define OFFSET_BITS 8
#define PRE_BITS 7
#define LITLEN_ENOUGH 2342
#define OFFSET_ENOUGH 402

/* entry: bits 0-3 length to consume (this level), 4-7 extra bits, 8-9 kind, 16-31 value */
enum { E_LITERAL = 0, E_BASE = 1, E_EOB = 2, E_SUBTABLE = 3 };
static uint32_t mk(uint32_t kind, uint32_t value, uint32_t extra, uint32_t len) { return len | (extra << 4) | (kind << 8) | (value << 16); }
#define E_LEN(e) ((e) & 15u)
#define E_EXTRA(e) (((e) >> 4) & 15u)
#define E_KIND(e) (((e) >> 8) & 3u)
#define E_VALUE(e) ((e) >> 16)

static const uint16_t k_len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t k_len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t k_off_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t k_off_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

static uint32_t result_for(int which, uint32_t sym, uint32_t len)
{
    if (which == 0) return mk(E_LITERAL, sym, 0, len);           /* precode: the symbol itself */
    if (which == 1) {                                             /* litlen */
        if (sym < 256) return mk(E_LITERAL, sym, 0, len);
        if (sym == 256) return mk(E_EOB, 0, 0, len);
        if (sym > 285) sym = 285;   /* 286/287 cannot occur in valid data; map like the reference's table does */
        return mk(E_BASE, k_len_base[sym - 257], k_len_extra[sym - 257], len);
    }
    if (sym > 29) sym = 29;
    return mk(E_BASE, k_off_base[sym], k_off_extra[sym], len);
}

/* canonical code -> direct table with one level of sub-tables; validity rules of deflate_decompress.c:799-853 */
static int build_table(uint32_t *table, const uint8_t *lens, unsigned num_syms, int which, unsigned table_bits, unsigned max_len,
                       unsigned *table_bits_ret)
{
    unsigned counts[16] = { 0 }, offsets[16];
    uint16_t sorted[288];
    for (unsigned s = 0; s < num_syms; s++) counts[lens[s]]++;
    while (max_len > 1 && counts[max_len] == 0) max_len--;
    if (table_bits_ret) {
        if (table_bits > max_len) table_bits = max_len;
        *table_bits_ret = table_bits;
    }
    offsets[0] = 0;
    offsets[1] = counts[0];
    uint32_t used = 0;
    unsigned len;
    for (len = 1; len < max_len; len++) {
        offsets[len + 1] = offsets[len] + counts[len];
        used = (used << 1) + counts[len];
    }
    used = (used << 1) + counts[len];
    for (unsigned s = 0; s < num_syms; s++) sorted[offsets[lens[s]]++] = (uint16_t)s;
    const uint16_t *syms = sorted + counts[0];
    if (used > (1u << max_len)) return 0;
    if (used < (1u << max_len)) {
        unsigned sym;
        if (used == 0) sym = 0;
        else {
            if (used != (1u << (max_len - 1)) || counts[1] != 1) return 0;
            sym = syms[0];
        }
        for (unsigned i = 0; i < (1u << table_bits); i++) table[i] = result_for(which, sym, 1);
        return 1;
    }
    /* complete code: walk codewords in canonical order */
    uint32_t code = 0;               /* canonical (MSB-first) codeword */
    unsigned next_sub = 1u << table_bits, si = 0;
    int cur_prefix = -1;
    unsigned cur_start = 0, cur_bits = 0;
    for (len = 1; len <= max_len; len++) {
        for (unsigned k = 0; k < counts[len]; k++, si++, code++) {
            uint32_t rev = 0;
            for (unsigned b = 0; b < len; b++) rev |= ((code >> b) & 1u) << (len - 1 - b);
            unsigned sym = syms[si];
            if (len <= table_bits) {
                for (uint32_t i = rev; i < (1u << table_bits); i += 1u << len) table[i] = result_for(which, sym, len);
            } else {
                uint32_t prefix = rev & ((1u << table_bits) - 1);
                if ((int)prefix != cur_prefix) {
                    /* new sub-table: size it for the longest codeword that shares this prefix */
                    cur_prefix = (int)prefix;
                    cur_start = next_sub;
                    cur_bits = len - table_bits;
                    uint32_t space = counts[len] - k;    /* codewords of this length still to place */
                    unsigned l2 = len;
                    while (space < (1u << cur_bits)) {
                        cur_bits++;
                        l2++;
                        space = (space << 1) + counts[l2];
                    }
                    next_sub += 1u << cur_bits;
                    table[prefix] = mk(E_SUBTABLE, cur_start, cur_bits, table_bits);
                }
                unsigned sublen = len - table_bits;
                for (uint32_t i = rev >> table_bits; i < (1u << cur_bits); i += 1u << sublen)
                    table[cur_start + i] = result_for(which, sym, sublen);
            }
        }
        code <<= 1;
    }
    return 1;
}

typedef struct {
    const uint8_t *p, *end;
    uint64_t bits;
    unsigned n;
    unsigned overread;
} bitsrc;

static void fill(bitsrc *b)
{
    while (b->n <= 56) {
        uint64_t byte = 0;
        if (b->p < b->end) byte = *b->p++;
        else b->overread++;
        b->bits |= byte << b->n;
        b->n += 8;
    }
}
static uint32_t take(bitsrc *b, unsigned k)
{
    uint32_t v = (uint32_t)(b->bits & ((1ull << k) - 1));
    b->bits >>= k;
    b->n -= k;
    return v;
}

static uint32_t decode_sym(bitsrc *b, const uint32_t *table, unsigned table_bits)
{
    uint32_t e = table[b->bits & ((1u << table_bits) - 1)];
    if (E_KIND(e) == E_SUBTABLE) {
        take(b, E_LEN(e));
        e = table[E_VALUE(e) + (b->bits & ((1u << E_EXTRA(e)) - 1))];
    }
    take(b, E_LEN(e));
    return e;
}

int oracle_inflate(uint8_t *out, size_t *out_len, const uint8_t *in, size_t in_len)
{
    static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    uint32_t pre[1 << PRE_BITS], *lit = malloc(sizeof(uint32_t) * LITLEN_ENOUGH), *off = malloc(sizeof(uint32_t) * OFFSET_ENOUGH);
    uint8_t lens[288 + 32 + 138];
    bitsrc b = { in, in + in_len, 0, 0, 0 };
    size_t cap = *out_len, pos = 0;
    int ret = OR_OK, final;
    do {
        fill(&b);
        final = (int)take(&b, 1);
        unsigned type = take(&b, 2);
        unsigned lit_bits = LITLEN_BITS;
        if (type == 0) {
            take(&b, b.n & 7);
            fill(&b);
            uint32_t len = take(&b, 16), nlen = take(&b, 16);
            if (len != (uint16_t)~nlen) { ret = OR_BAD_DATA; break; }
            /* un-read the real bytes still sitting in the bit buffer (zero padding read past the end is not real) */
            if (b.overread > (b.n >> 3)) { ret = OR_BAD_DATA; break; }
            const uint8_t *raw = b.p - ((b.n >> 3) - b.overread);
            if (len > (size_t)(b.end - raw)) { ret = OR_BAD_DATA; break; }
            if (len > cap - pos) { ret = OR_INSUFFICIENT_SPACE; break; }
            memcpy(out + pos, raw, len);
            pos += len;
            b.p = raw + len;
            b.bits = 0;
            b.n = 0;
            b.overread = 0;
            continue;
        } else if (type == 1) {
            unsigned i;
            for (i = 0; i < 144; i++) lens[i] = 8;
            for (; i < 256; i++) lens[i] = 9;
            for (; i < 280; i++) lens[i] = 7;
            for (; i < 288; i++) lens[i] = 8;
            for (; i < 288 + 32; i++) lens[i] = 5;
            if (!build_table(off, lens + 288, 32, 2, OFFSET_BITS, 15, NULL) || !build_table(lit, lens, 288, 1, LITLEN_BITS, 15, &lit_bits)) { ret = OR_BAD_DATA; break; }
        } else if (type == 2) {
            unsigned nl = take(&b, 5) + 257, nd = take(&b, 5) + 1, np = take(&b, 4) + 4, i;
            uint8_t plens[19] = { 0 };
            for (i = 0; i < np; i++) { fill(&b); plens[order[i]] = (uint8_t)take(&b, 3); }
            if (!build_table(pre, plens, 19, 0, PRE_BITS, 7, NULL)) { ret = OR_BAD_DATA; break; }
            i = 0;
            while (i < nl + nd) {
                fill(&b);
                uint32_t e = decode_sym(&b, pre, PRE_BITS);
                unsigned sym = E_VALUE(e);
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                unsigned rep;
                uint8_t v = 0;
                if (sym == 16) {
                    if (i == 0) { ret = OR_BAD_DATA; break; }
                    v = lens[i - 1];
                    rep = 3 + take(&b, 2);
                } else if (sym == 17) rep = 3 + take(&b, 3);
                else rep = 11 + take(&b, 7);
                /* the reference lets a run overshoot into slack and checks afterwards (decompress_template.h:194-236) */
                for (unsigned k = 0; k < rep; k++) lens[i + k] = v;
                i += rep;
            }
            if (ret) break;
            if (i != nl + nd) { ret = OR_BAD_DATA; break; }
            if (!build_table(off, lens + nl, nd, 2, OFFSET_BITS, 15, NULL) || !build_table(lit, lens, nl, 1, LITLEN_BITS, 15, &lit_bits)) { ret = OR_BAD_DATA; break; }
        } else { ret = OR_BAD_DATA; break; }
        for (;;) {
            fill(&b);
            uint32_t e = decode_sym(&b, lit, lit_bits);
            unsigned kind = E_KIND(e);
            if (kind == E_LITERAL) {
                if (pos >= cap) { ret = OR_INSUFFICIENT_SPACE; break; }
                out[pos++] = (uint8_t)E_VALUE(e);
                continue;
            }
            if (kind == E_EOB) break;
            uint32_t len = E_VALUE(e) + take(&b, E_EXTRA(e));
            fill(&b);
            e = decode_sym(&b, off, OFFSET_BITS);
            uint32_t dist = E_VALUE(e) + take(&b, E_EXTRA(e));
            if (dist > pos) { ret = OR_BAD_DATA; break; }
            if (len > cap - pos) { ret = OR_INSUFFICIENT_SPACE; break; }
            for (uint32_t k = 0; k < len; k++, pos++) out[pos] = out[pos - dist];
            if (b.overread > 8) { ret = OR_BAD_DATA; break; }
        }
        if (ret) break;
        if (b.overread > 8) { ret = OR_BAD_DATA; break; }
    } while (!final);
    /* input that ended inside the bit buffer's zero padding was truncated */
    if (!ret && b.overread > (b.n >> 3)) ret = OR_BAD_DATA;
    /* like lib/zlibutil.c:199 (actual_out_nbytes_ret != NULL): a short output is not an error, *out_len reports it */
    *out_len = pos;
    free(lit);
    free(off);
    return ret;
}

/* applet/7bgzf.c:295-365: walk members, inflate each into ISIZE bytes.  Returns 0, -1 (not BGZF) or 1 (inflate error). */
int oracle_bgzf_decompress(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, size_t *out_len, size_t *members)
{
    size_t ipos = 0, opos = 0, nm = 0;
    while (ipos < in_len) {
        int extra_off, extra_len, block_len = 0;
        size_t avail = in_len - ipos;
        int n = oracle_read_gz_header(in + ipos, avail > 64 ? 64 : (int)avail, &extra_off, &extra_len, &block_len);
        if (!n || block_len < n + 8 || (size_t)block_len > avail) return -1;
        size_t isize = get32(in + ipos + block_len - 4);
        if (isize > out_cap - opos) return 1;
        size_t got = isize;
        if (oracle_inflate(out + opos, &got, in + ipos + n, (size_t)block_len - n - 8) != OR_OK) return 1;
        opos += got;
        ipos += block_len;
        nm++;
    }
    *out_len = opos;
    if (members) *members = nm;
    return 0;
}

/* bgzf_compress.c:53-113: returns the method enum of lib/zlibutil.h:13-26 and the level the hook would use */
int oracle_parse_method(const char *spec, int *level_out)
{
    enum { ZLIB = 0, SEVENZIP, ZOPFLI, MINIZ, SLZ, LIBDEFLATE, ZLIBNG, IGZIP, CRYPTOPP };
    int method = ZLIB, level = -1;
    if (spec && *spec) {
        char s[256];
        strncpy(s, spec, sizeof s - 1);
        s[sizeof s - 1] = 0;
        int l = (int)strlen(s), i = l - 1, digit = 1, lv = -1;
        for (; i >= 0 && s[i] >= '0' && s[i] <= '9'; i--) {
            if (lv < 0) lv = 0;
            lv += digit * (s[i] - '0');
            digit *= 10;
        }
        if (lv >= 0) level = lv;
        s[i + 1] = 0;
        if (!strcasecmp(s, "zlib")) method = ZLIB;
        if (!strcasecmp(s, "7zip") || !strcasecmp(s, "7-zip")) method = SEVENZIP;
        if (!strcasecmp(s, "zopfli")) method = ZOPFLI;
        if (!strcasecmp(s, "miniz")) method = MINIZ;
        if (!strcasecmp(s, "slz") || !strcasecmp(s, "libslz")) method = SLZ;
        if (!strcasecmp(s, "libdeflate")) method = LIBDEFLATE;
        if (!strcasecmp(s, "zlibng")) method = ZLIBNG;
        if (!strcasecmp(s, "igzip")) method = IGZIP;
        if (!strcasecmp(s, "cryptopp")) method = CRYPTOPP;
    }
    if (level < 0) {
        static const int defaults[] = { 6, 2, 1, 1, 1, 6, 6, 1, 6 };
        level = defaults[method];
    }
    *level_out = level;
    return method;
}
