// Throw-away ratio experiment: all-position hash-chain match search + local lazy parse + one dynamic
// Huffman block per BGZF block.  Prints the total BGZF size for a file so that design parameters
// (hash bits, chain depth, nice length, lazy rule, min match length) can be compared with the reference's
// libdeflate sizes.  Not part of the product or the tests.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static int HB = 14, DEPTH = 32, NICE = 65, HBYTES = 4, PARSE = 1, MINLEN_MODE = 1, PASSES = 1, OPT = 0, NCAND = 1;

static const uint16_t len_base[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t len_extra[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t off_base[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t off_extra[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static uint8_t len_slot[259];
static int off_slot(unsigned off) {
    int s = 0;
    for (int i = 29; i >= 0; i--) if (off >= off_base[i]) { s = i; break; }
    return s;
}
static uint8_t off_slot_tab[32769];

// Huffman code lengths, limited to maxbits.  freq[n] -> lens[n]
static void huff_lengths(const uint32_t *freq, int n, int maxbits, uint8_t *lens) {
    struct Node { uint64_t w; int sym; int l, r; };
    std::vector<std::pair<uint32_t,int>> used;
    for (int i = 0; i < n; i++) { lens[i] = 0; if (freq[i]) used.push_back({freq[i], i}); }
    if (used.empty()) return;
    if (used.size() == 1) { lens[used[0].second] = 1; return; }
    std::sort(used.begin(), used.end());
    int m = used.size();
    std::vector<uint64_t> w(2*m); std::vector<int> parent(2*m, -1);
    for (int i = 0; i < m; i++) w[i] = used[i].first;
    int a = 0, b = m, e = m; // leaf queue [a,m), internal queue [b,e)
    while ((m - a) + (e - b) > 1) {
        int x[2];
        for (int k = 0; k < 2; k++) {
            if (a < m && (b >= e || w[a] <= w[b])) x[k] = a++; else x[k] = b++;
        }
        w[e] = w[x[0]] + w[x[1]]; parent[x[0]] = e; parent[x[1]] = e; e++;
    }
    std::vector<int> depth(2*m, 0);
    for (int i = e - 2; i >= 0; i--) depth[i] = depth[parent[i]] + 1;
    // length limiting: clamp then fix Kraft
    std::vector<int> L(m);
    for (int i = 0; i < m; i++) L[i] = std::min(depth[i], maxbits);
    uint64_t kraft = 0; // in units of 2^-maxbits
    for (int i = 0; i < m; i++) kraft += 1ull << (maxbits - L[i]);
    uint64_t one = 1ull << maxbits;
    // over-subscribed: lengthen the least frequent symbols that are shorter than maxbits (from the rarest)
    while (kraft > one) {
        for (int i = 0; i < m && kraft > one; i++) {
            if (L[i] < maxbits) { kraft -= 1ull << (maxbits - L[i] - 1); L[i]++; }
        }
    }
    // under-subscribed: shorten the most frequent where possible
    for (int i = m - 1; i >= 0; i--) {
        while (L[i] > 1 && kraft + (1ull << (maxbits - L[i])) <= one) { kraft += 1ull << (maxbits - L[i]); L[i]--; }
    }
    for (int i = 0; i < m; i++) lens[used[i].second] = L[i];
}

static const uint8_t pre_order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};

// dynamic header bit cost
static uint32_t header_bits(const uint8_t *ll, const uint8_t *dl) {
    int nl = 286; while (nl > 257 && ll[nl-1] == 0) nl--;
    int nd = 30; while (nd > 1 && dl[nd-1] == 0) nd--;
    uint8_t lens[320]; memcpy(lens, ll, nl); memcpy(lens + nl, dl, nd);
    int n = nl + nd;
    uint32_t pf[19] = {0}; uint32_t extra = 0;
    for (int i = 0; i < n;) {
        int j = i; while (j < n && lens[j] == lens[i]) j++;
        int run = j - i; uint8_t v = lens[i];
        if (v == 0) {
            while (run >= 11) { int r = std::min(run, 138); pf[18]++; extra += 7; run -= r; }
            if (run >= 3) { pf[17]++; extra += 3; run = 0; }
            pf[0] += run;
        } else {
            pf[v]++; run--;
            while (run >= 3) { int r = std::min(run, 6); pf[16]++; extra += 2; run -= r; }
            pf[v] += run;
        }
        i = j;
    }
    uint8_t pl[19]; huff_lengths(pf, 19, 7, pl);
    int np = 19; while (np > 4 && pl[pre_order[np-1]] == 0) np--;
    uint32_t bits = 3 + 5 + 5 + 4 + 3 * np + extra;
    for (int i = 0; i < 19; i++) bits += pf[i] * pl[i];
    return bits;
}

struct Tok { uint16_t pos, len, off; };

static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline int bsr(uint32_t v) { return 31 - __builtin_clz(v); }

static size_t compress_block(const uint8_t *in, int n, long *stats) {
    static std::vector<uint16_t> head, prev; static std::vector<uint16_t> mlen, moff;
    head.assign(1 << HB, 0xffff); prev.assign(n + 8, 0xffff); mlen.assign(n + 8, 0); moff.assign(n + 8, 0);
    static std::vector<uint16_t> mlen2[3], moff2[3]; for (int c = 0; c < 3; c++) { mlen2[c].assign(n + 8, 0); moff2[c].assign(n + 8, 0); }
    std::vector<uint8_t> buf(n + 300, 0); memcpy(buf.data(), in, n); in = buf.data();
    uint64_t hmask = HBYTES >= 8 ? ~0ull : ((1ull << (8 * HBYTES)) - 1);
    auto hash = [&](int p) { uint64_t v; memcpy(&v, in + p, 8); v &= hmask;
        if (HBYTES <= 4) return (uint32_t)(((uint32_t)v * 0x1E35A7BDu) >> (32 - HB));
        if (HBYTES == 8) return (uint32_t)((((uint32_t)v * 0x1E35A7BDu) ^ ((uint32_t)(v >> 32) * 0x9E3779B1u)) >> (32 - HB));
        return (uint32_t)((v * 0x9E3779B185EBCA87ull) >> (64 - HB)); };
    int last = n - HBYTES; // positions with a full hash
    for (int p = 0; p <= last; p++) { uint32_t h = hash(p); prev[p] = head[h]; head[h] = p; }
    // all-position search
    for (int p = 0; p < n; p++) {
        int best = 0, boff = 0; int maxl = std::min(258, n - p);
        if (p <= last && maxl >= 3) {
            int q = prev[p]; int d = DEPTH;
            while (q != 0xffff && d-- > 0) {
                int dist = p - q; if (dist > 32768) break;
                if (rd32(in + q) == rd32(in + p) || (best < 3 && (rd32(in+q) & 0xffffff) == (rd32(in+p) & 0xffffff))) {
                    if (best < 3 || in[q + best] == in[p + best]) {
                        int l = 0; while (l < maxl && in[q + l] == in[p + l]) l++;
                        if (l > best) { if (best >= 3) { for (int c = 2; c > 0; c--) { mlen2[c][p] = mlen2[c-1][p]; moff2[c][p] = moff2[c-1][p]; } mlen2[0][p] = best; moff2[0][p] = boff; } best = l; boff = dist; if (l >= NICE) break; }
                    }
                }
                q = prev[q];
            }
        }
        if (best >= 3 && best <= maxl) { mlen[p] = best; moff[p] = boff; }
    }
    // min match len from distinct literals in block
    int minlen = 3;
    if (MINLEN_MODE && n >= 512) {
        bool used[256] = {0}; int nu = 0; for (int i = 0; i < n; i++) used[in[i]] = true; for (int i = 0; i < 256; i++) nu += used[i];
        static const uint8_t ml[] = {9,9,9,9,9,9,8,8,7,7,6,6,6,6,6,6,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4};
        minlen = nu < (int)sizeof(ml) ? ml[nu] : 3;
        minlen += MINLEN_MODE - 1;
    }
    uint8_t ll[288], dl[32]; bool have_costs = false;
    uint32_t bits = 0;
    for (int pass = 0; pass < PASSES; pass++) {
        std::vector<Tok> toks;
        auto ok = [&](int p) {
            int l = mlen[p]; if (l < 3) return false;
            if (!have_costs) { if (l < minlen) return false; if (l == 3 && moff[p] > 8192) return false; return true; }
            // cost-based: match cost vs literal cost
            int ls = len_slot[l], os = off_slot_tab[moff[p]];
            int mc = (ll[257+ls] ? ll[257+ls] : 15) + len_extra[ls] + (dl[os] ? dl[os] : 15) + off_extra[os];
            int lc = 0; for (int k = 0; k < l; k++) lc += ll[in[p+k]] ? ll[in[p+k]] : 15;
            return mc < lc;
        };
        for (int p = 0; p < n;) {
            if (!ok(p)) { toks.push_back({(uint16_t)p, 0, 0}); p++; continue; }
            int cl = mlen[p], co = moff[p];
            if (PARSE >= 1) {
                while (cl < NICE && p + 1 < n && ok(p + 1)) {
                    int nl = mlen[p+1], no = moff[p+1];
                    if (nl >= cl && 4 * (nl - cl) + (bsr(co) - bsr(no)) > 2) { toks.push_back({(uint16_t)p, 0, 0}); p++; cl = nl; co = no; continue; }
                    if (PARSE >= 2 && p + 2 < n && ok(p + 2)) {
                        int nl2 = mlen[p+2], no2 = moff[p+2];
                        if (nl2 >= cl && 4 * (nl2 - cl) + (bsr(co) - bsr(no2)) > 6) { toks.push_back({(uint16_t)p, 0, 0}); toks.push_back({(uint16_t)(p+1), 0, 0}); p += 2; cl = nl2; co = no2; continue; }
                    }
                    break;
                }
            }
            toks.push_back({(uint16_t)p, (uint16_t)cl, (uint16_t)co}); p += cl;
        }
        uint32_t lf[288] = {0}, df[32] = {0};
        for (auto &t : toks) { if (t.len) { lf[257 + len_slot[t.len]]++; df[off_slot_tab[t.off]]++; } else lf[in[t.pos]]++; }
        lf[256] = 1;
        huff_lengths(lf, 288, 15, ll); huff_lengths(df, 32, 15, dl);
        int nd = 0; for (int i = 0; i < 32; i++) nd += dl[i] != 0; if (nd == 0) dl[0] = 1;
        bits = header_bits(ll, dl);
        for (int i = 0; i < 288; i++) bits += lf[i] * ll[i];
        for (int i = 0; i < 30; i++) bits += df[i] * dl[i];
        for (int i = 0; i < 29; i++) bits += lf[257 + i] * len_extra[i];
        for (int i = 0; i < 30; i++) bits += df[i] * off_extra[i];
        have_costs = true;
        if (pass == PASSES - 1) { for (auto &t : toks) { if (t.len) { stats[1]++; stats[2] += t.len; } else stats[0]++; } }
    }
    for (int it = 0; it < OPT; it++) {
        // costs from current ll/dl (0 -> 13 bits guess)
        std::vector<uint32_t> cost(n + 1, 0); std::vector<uint16_t> choice(n + 1, 1);
        auto lc = [&](int sym) { return ll[sym] ? ll[sym] : 13; };
        auto dc = [&](int s2) { return dl[s2] ? dl[s2] : 10; };
        for (int p = n - 1; p >= 0; p--) {
            uint32_t best = lc(in[p]) + cost[p + 1]; int bl = 1;
            for (int c = 0; c < NCAND; c++) {
                int L = c == 0 ? mlen[p] : mlen2[c-1][p], O = c == 0 ? moff[p] : moff2[c-1][p];
                if (L < 3) continue;
                uint32_t oc = dc(off_slot_tab[O]) + off_extra[off_slot_tab[O]];
                for (int l = 3; l <= L; l++) {
                    uint32_t v = lc(257 + len_slot[l]) + len_extra[len_slot[l]] + oc + cost[p + l];
                    if (v < best) { best = v; bl = l | (c << 12); }
                }
            }
            cost[p] = best; choice[p] = bl;
        }
        uint32_t lf[288] = {0}, df[32] = {0};
        for (int p = 0; p < n;) { int bl = choice[p] & 0xfff, c = choice[p] >> 12; if (bl == 1) { lf[in[p]]++; p++; } else { int O = c == 0 ? moff[p] : moff2[c-1][p]; lf[257 + len_slot[bl]]++; df[off_slot_tab[O]]++; p += bl; } }
        lf[256] = 1;
        huff_lengths(lf, 288, 15, ll); huff_lengths(df, 32, 15, dl);
        bits = header_bits(ll, dl);
        for (int i = 0; i < 288; i++) bits += lf[i] * ll[i];
        for (int i = 0; i < 30; i++) bits += df[i] * dl[i];
        for (int i = 0; i < 29; i++) bits += lf[257 + i] * len_extra[i];
        for (int i = 0; i < 30; i++) bits += df[i] * off_extra[i];
    }
    size_t bytes = (bits + 7) / 8;
    size_t stored = n + 5;
    return 26 + std::min(bytes, stored);
}

int main(int argc, char **argv) {
    const char *path = argv[1]; size_t limit = strtoull(argv[2], 0, 10);
    for (int i = 3; i < argc; i++) {
        if (!strncmp(argv[i], "hb=", 3)) HB = atoi(argv[i] + 3);
        if (!strncmp(argv[i], "depth=", 6)) DEPTH = atoi(argv[i] + 6);
        if (!strncmp(argv[i], "nice=", 5)) NICE = atoi(argv[i] + 5);
        if (!strncmp(argv[i], "hbytes=", 7)) HBYTES = atoi(argv[i] + 7);
        if (!strncmp(argv[i], "parse=", 6)) PARSE = atoi(argv[i] + 6);
        if (!strncmp(argv[i], "minlen=", 7)) MINLEN_MODE = atoi(argv[i] + 7);
        if (!strncmp(argv[i], "passes=", 7)) PASSES = atoi(argv[i] + 7);
        if (!strncmp(argv[i], "opt=", 4)) OPT = atoi(argv[i] + 4);
        if (!strncmp(argv[i], "ncand=", 6)) NCAND = atoi(argv[i] + 6);
    }
    for (int l = 3; l <= 258; l++) { int s = 0; for (int i = 28; i >= 0; i--) if (l >= len_base[i]) { s = i; break; } len_slot[l] = s; }
    for (int o = 1; o <= 32768; o++) off_slot_tab[o] = off_slot(o);
    FILE *f = fopen(path, "rb"); std::vector<uint8_t> data(limit); size_t n = fread(data.data(), 1, limit, f); fclose(f);
    size_t total = 0; long stats[3] = {0,0,0};
    for (size_t o = 0; o < n; o += 0xff00) total += compress_block(data.data() + o, (int)std::min<size_t>(0xff00, n - o), stats);
    printf("hb=%d hbytes=%d depth=%d nice=%d parse=%d minlen=%d passes=%d : out=%zu ratio=%.4f  lits=%ld matches=%ld avglen=%.1f\n",
           HB, HBYTES, DEPTH, NICE, PARSE, MINLEN_MODE, PASSES, total, (double)total / n, stats[0], stats[1], stats[1] ? (double)stats[2]/stats[1] : 0.0);
}
