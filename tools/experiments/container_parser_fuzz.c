/* Mutation fuzz of the container index parsers (7bgzf_b200/host/containers.c: b200bgzf_container_units) under ASAN + UBSAN, no GPU:
 *   gcc -O1 -g -fsanitize=address,undefined -Iinclude -o /tmp/cpf tools/experiments/container_parser_fuzz.c 7bgzf_b200/host/containers.c
 *   G=tests/golden/containers; /tmp/cpf 4 $G/ref.dz 5 $G/ref.raz 3 $G/ref.gzinga 1 $G/ref.gz
 * 20000 mutants per file (bit flips and random bytes in the header, the trailing index and anywhere; random truncation), each
 * parsed from an exact-size heap block; every accepted unit must lie inside the input. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/b200bgzf.h"
/* stubs for the GPU entry points containers.c refers to */
size_t b200bgzf_compress_bound(size_t n, uint32_t bs) { return n + 38 * ((n + bs - 1) / bs) + 28; }
size_t b200bgzf_pieces_gap_bytes(size_t a, uint32_t b, const b200bgzf_piece_spec *s) { return 0; }
int b200bgzf_compress_pieces_host(b200bgzf_ctx *c, const void *in, size_t n, uint32_t bs, int l, const b200bgzf_piece_spec *s, void *o, size_t cap, size_t *ob, uint64_t *po, uint32_t *pc, size_t pcap) { return -2; }
int b200bgzf_inflate_host(b200bgzf_ctx *c, const void *in, size_t n, void *o, size_t cap, size_t *ob, unsigned f) { return -2; }
int b200bgzf_inflate_units_host(b200bgzf_ctx *c, const void *in, size_t n, const b200bgzf_unit *u, size_t nu, void *o, size_t cap, size_t *ob, unsigned f, uint32_t *crc) { return -2; }
int b200bgzf_inflate_size_host(const void *in, size_t n, size_t *ob, size_t *nm) { return -3; }
static unsigned long long rs = 88172645463325252ull;
static unsigned rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)(rs >> 11); }
int main(int argc, char **argv)
{
    long ok = 0, bad = 0;
    for (int f = 1; f + 1 < argc; f += 2) {
        int kind = atoi(argv[f]);
        FILE *fp = fopen(argv[f + 1], "rb");
        fseek(fp, 0, SEEK_END); long n = ftell(fp); fseek(fp, 0, SEEK_SET);
        unsigned char *orig = malloc(n), *buf = malloc(n);
        fread(orig, 1, n, fp); fclose(fp);
        for (int it = 0; it < 20000; it++) {
            long len = n;
            unsigned char *m = malloc(n);      /* exact-size heap block: ASAN catches any read past it */
            memcpy(m, orig, n);
            int nm = 1 + rnd() % 4;
            for (int k = 0; k < nm; k++) {
                /* mutate mostly in the header and in the trailing index, where the parsers look */
                long pos = rnd() % 3 == 0 ? rnd() % n : rnd() % 2 ? rnd() % (n < 200 ? n : 200) : n - 1 - rnd() % (n < 400 ? n : 400);
                m[pos] = rnd() % 4 == 0 ? rnd() : m[pos] ^ (1u << (rnd() % 8));
            }
            if (rnd() % 5 == 0) len = rnd() % (n + 1);
            unsigned char *t = malloc(len ? len : 1);
            memcpy(t, m, len);
            free(m);
            b200bgzf_unit *u = NULL; size_t nu = 0, ob = 0;
            int rc = b200bgzf_container_units(kind, t, len, &u, &nu, &ob);
            if (rc == 0) {
                ok++;
                for (size_t i = 0; i < nu; i++)
                    if (u[i].in_off + u[i].in_len > (size_t)len || u[i].hdr_len > u[i].in_len) { printf("unit out of range kind %d\n", kind); return 1; }
                b200bgzf_units_free(u);
            } else bad++;
            free(t);
        }
        free(orig); free(buf);
    }
    printf("accepted %ld rejected %ld\n", ok, bad);
    return 0;
}
