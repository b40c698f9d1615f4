/*
 * bgzf_tables.h — host-side generators for the two constant tables the block kernels read:
 * the byte-wise CRC-32 table and the slice-combine multipliers x^(8*68*k) mod P
 * (the role crc32_combine_gen plays in lib/zlib/crc32.c:1021-1049 of the reference).
 */
#ifndef BGZF_TABLES_H
#define BGZF_TABLES_H
#include "bgzf_block.h"

static inline void bg_make_crc_table(uint32_t tab[256])
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t r = i;
        for (int k = 0; k < 8; k++)
            r = (r >> 1) ^ (BG_CRC_POLY & (0u - (r & 1u)));
        tab[i] = r;
    }
}

/* x^nbits mod P, reflected representation (x^0 = 0x80000000) */
static inline uint32_t bg_crc_xpow(uint64_t nbits)
{
    uint32_t result = 0x80000000u, sq = 0x40000000u; /* x^1 */
    while (nbits) {
        if (nbits & 1) result = bg_crc_mul(result, sq);
        sq = bg_crc_mul(sq, sq);
        nbits >>= 1;
    }
    return result;
}

static inline void bg_make_crc_pow(uint32_t *pow, uint32_t count)
{
    const uint32_t base = bg_crc_xpow(8ull * 4 * BG_CRC_WORDS);
    pow[0] = 0x80000000u;
    for (uint32_t k = 1; k < count; k++)
        pow[k] = bg_crc_mul(pow[k - 1], base);
}
#endif
