/*
 * bgzf_compress.cu — sm_100a BGZF compress kernels.
 *
 *   bgzf_compress_kernel : persistent, one CTA (1024 threads) per SM; each CTA takes BGZF blocks b, b+grid, ...
 *       1. the <=64 KiB payload is staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier)
 *       2. CRC-32 slices + literal census, hash of every position             (bgzf_block.h phases)
 *       3. hash chains linked by warp 0 with __match_any_sync (in position order => deterministic)
 *       4. all-position chain search, 1024 positions at a time, results to an L2-resident scratch
 *       5. local lazy rule -> step codes, per-chunk jump table, chunk-to-chunk walk
 *       6. histograms, length-limited Huffman (bitonic sort + two-queue tree), block type choice
 *       7. per-chunk bit sizes, block-wide prefix sum, parallel bit packing, BGZF header/BSIZE/CRC/ISIZE
 *   bgzf_scan_kernel / bgzf_gather_kernel : exclusive scan of member sizes and compaction of the
 *       fixed-stride slots into one contiguous BGZF stream (+ the 28-byte EOF marker).
 *
 * Shared memory map (bytes): [0,65568) payload | [65568,196640) region A | [196640,229408) region B |
 * crc table 1 KiB | literal flags | scalars | mbarrier | scan scratch  = 230,968 B of the 232,448 B a CTA may own.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "bgzf_block.h"
#include "bgzf_kernels.h"

#define SM_DATA 0u
#define SM_REGA (BG_DATA_BYTES)
#define SM_REGB (SM_REGA + 131072u)
#define SM_CRCTAB (SM_REGB + 32768u)
#define SM_LITFLAG (SM_CRCTAB + 1024u)
#define SM_SCAL (SM_LITFLAG + 256u)
#define SM_MBAR (SM_SCAL + 4u * BG_S_COUNT)
#define SM_FRONTIER (SM_MBAR + 8u)
#define SM_SCAN (SM_MBAR + 16u)
#define SM_TOTAL (SM_SCAN + 4u * 36u)

extern "C" size_t bgzf_compress_smem_bytes(void) { return SM_TOTAL; }

/* ---- TMA bulk copy + mbarrier (PTX ISA: cp.async.bulk, mbarrier) ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/* ---- one warp links the hash chains, 32 positions per step, in position order, and publishes how far it got.
 * Lanes holding the same hash inside a step are found with one ballot per hash bit (constant time; the
 * match.any instruction serialises over distinct values and cost 400+ cycles per step here). ---- */
__device__ __forceinline__ void build_chains_warp(const BgCtx &c, uint32_t lane, volatile uint32_t *frontier)
{
    const uint32_t n = c.n;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t hnext = lane < n ? c.prev[lane] : BG_NOPOS;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t h = hnext;
        hnext = p + 32 < n ? c.prev[p + 32] : BG_NOPOS;      /* next step's hashes: those slots are still hashes */
        const bool valid = h != BG_NOPOS;
        unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int k = 0; k < BG_HASH_BITS; k++) {
            const bool bit = (h >> k) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? b : ~b;
        }
        if (valid) {
            const unsigned lower = peers & lt;
            const uint32_t link = lower ? base + (31u - (uint32_t)__clz(lower)) : (uint32_t)c.head[h];
            c.prev[p] = (uint16_t)link;
            if ((peers >> lane) == 1u)       /* highest lane holding this hash */
                c.head[h] = (uint16_t)p;
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            *frontier = base + 32;           /* links of every position below this are final */
        }
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        *frontier = 0xffffffffu;
    }
}

/* ---- all-position search: every thread owns positions t, t+1024, ... and works through them with the
 * state machine of bgzf_block.h, one 4-byte comparison per iteration, so a warp never waits for its slowest
 * lane except at the very end.  A position is started only once the builder has linked everything up to it. ---- */
__device__ __forceinline__ void search_positions(const BgCtx &c, uint32_t t, volatile const uint32_t *frontier)
{
    const uint32_t n = c.n;
    uint32_t p = t, fr = 0;
    bool have = false, exhausted = p >= n;
    BgSearch s;
    s.p = s.q = s.maxl = s.best = s.boff = s.l = s.ptail = 0;
    s.depth = 0;
    s.ext = false;
    for (;;) {
        if (!have && !exhausted) {
            if (fr <= p) fr = *frontier;
            if (fr > p) {
                if (bg_search_begin(c, s, p)) {
                    have = true;
                } else {
                    c.R[p] = 0;
                    p += BG_THREADS;
                    exhausted = p >= n;
                }
            }
        }
        if (have && bg_search_step(c, s)) {
            c.R[p] = bg_search_result(s);
            have = false;
            p += BG_THREADS;
            exhausted = p >= n;
        }
        if (__all_sync(0xffffffffu, exhausted && !have)) break;
    }
}

/* ---- 512-key bitonic sort in shared memory (ascending); all threads call it ---- */
__device__ __forceinline__ void bitonic_sort_512(uint32_t *keys, uint32_t t)
{
    for (uint32_t k = 2; k <= 512; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            if (t < 512) {
                uint32_t x = t ^ j;
                if (x > t) {
                    uint32_t a = keys[t], b = keys[x];
                    bool up = (t & k) == 0;
                    if ((a > b) == up) { keys[t] = b; keys[x] = a; }
                }
            }
            __syncthreads();
        }
    }
}

/* ---- exclusive prefix sum over 1024 u32 in shared memory; returns the grand total ---- */
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t *v, uint32_t *scratch, uint32_t t)
{
    const uint32_t lane = t & 31u, warp = t >> 5;
    uint32_t x = v[t], inc = x;
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = scratch[lane], winc = w;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (uint32_t)d) winc += y;
        }
        scratch[lane] = winc - w;
        if (lane == 31) scratch[32] = winc;
    }
    __syncthreads();
    v[t] = scratch[warp] + inc - x;
    uint32_t total = scratch[32];
    __syncthreads();
    return total;
}

#define PROF_MARK(i)                                                        \
    do {                                                                    \
        if (prof && t == 0) {                                               \
            long long now_ = clock64();                                     \
            atomicAdd((unsigned long long *)&prof[i], (unsigned long long)(now_ - tprev_)); \
            tprev_ = now_;                                                  \
        }                                                                   \
    } while (0)

__global__ void __launch_bounds__(BG_THREADS, 1)
bgzf_compress_kernel(BgzfCompressArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t t = threadIdx.x, T = BG_THREADS;
    uint64_t *mbar = (uint64_t *)(smem + SM_MBAR);
    uint32_t *scan_scratch = (uint32_t *)(smem + SM_SCAN);
    volatile uint32_t *frontier = (volatile uint32_t *)(smem + SM_FRONTIER);
    unsigned long long *prof = a.prof;
    long long tprev_ = prof ? clock64() : 0;

    BgCtx c;
    c.dataw = (uint32_t *)(smem + SM_DATA);
    c.prev = (uint16_t *)(smem + SM_REGA);
    c.stepcode = smem + SM_REGA;
    c.jump8 = smem + SM_REGA + 65536u;
    c.offarr = (uint16_t *)(smem + SM_REGA + 65536u);
    c.head = (uint16_t *)(smem + SM_REGB);
    c.regb = smem + SM_REGB;
    c.crctab = (uint32_t *)(smem + SM_CRCTAB);
    c.litflag = smem + SM_LITFLAG;
    c.scal = (uint32_t *)(smem + SM_SCAL);
    c.R = a.scratch + (size_t)blockIdx.x * BGZF_SCRATCH_WORDS;
    c.crcpow = a.crcpow;
    c.prm = a.prm;

    if (t < 256) c.crctab[t] = a.crctab[t];
    if (t == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;

    for (uint32_t b = blockIdx.x; b < a.nblocks; b += gridDim.x) {
        const uint8_t *src;
        uint32_t n;
        if (a.in_off) {
            src = a.in + a.in_off[b];
            n = a.in_len[b];
        } else {
            const uint64_t off = (uint64_t)b * a.block_size;
            src = a.in + off;
            const uint64_t rem = a.in_bytes - off;
            n = rem < a.block_size ? (uint32_t)rem : a.block_size;
        }
        c.n = n;
        c.out = (uint32_t *)(a.slots + (size_t)b * BG_SLOT_BYTES);

        /* 1. stage the payload: TMA for the 16-byte-aligned bulk, plain loads for the ragged rest */
        const bool aligned = (((uintptr_t)src) & 15u) == 0;
        const uint32_t bulk = aligned ? (n & ~15u) : 0u;
        if (bulk) {
            if (t == 0) {
                fence_proxy_async();
                mbar_expect_tx(mbar, bulk);
                tma_load_1d(smem + SM_DATA, src, bulk, mbar);
            }
        }
        for (uint32_t i = bulk + t; i < n; i += T)
            smem[SM_DATA + i] = src[i];
        bg_phase_init(c, t, T);
        if (bulk) {
            mbar_wait(mbar, parity);
            parity ^= 1u;
        }
        __syncthreads();
        PROF_MARK(0);

        bg_phase_scan(c, t, T);
        __syncthreads();
        bg_phase_count(c, t, T);
        __syncthreads();
        bg_phase_settle(c, t, T);
        __syncthreads();
        PROF_MARK(1);
        bg_phase_hash(c, t, T);
        __syncthreads();
        PROF_MARK(2);
        const long long tbuild_ = prof ? clock64() : 0;
        /* chain build (warp 31, which the scheduler favours) overlapped with the search (everyone) */
        if (t == 0) *frontier = 0;
        __syncthreads();
        if (t >= BG_THREADS - 32) {
            build_chains_warp(c, t & 31u, frontier);
            if (prof && t == BG_THREADS - 32) atomicAdd(&prof[10], (unsigned long long)(clock64() - tbuild_));
        }
        search_positions(c, t, frontier);
        __syncthreads();
        PROF_MARK(4);
        bg_phase_accept(c, t, T);
        __syncthreads();
        bg_phase_jump(c, t, T);
        __syncthreads();
        PROF_MARK(5);
        bg_phase_walk_a(c, t, T);
        bg_phase_clear_freq(c, t, T);
        __syncthreads();
        bg_phase_walk_b(c, t, T);
        __syncthreads();
        bg_phase_walk_c(c, t, T);
        __syncthreads();
        PROF_MARK(6);
        bg_phase_tally(c, t, T);
        __syncthreads();
        bg_phase_lkeys(c, t, T);
        __syncthreads();
        PROF_MARK(7);
        bitonic_sort_512((uint32_t *)(c.regb + BG_B_KEYS), t);
        bg_phase_huff(c, t, T);
        __syncthreads();
        bg_phase_decide(c, t, T);
        __syncthreads();
        PROF_MARK(8);
        bg_phase_sizes(c, t, T);
        __syncthreads();
        block_exclusive_scan_1024((uint32_t *)(c.regb + BG_B_CBITS), scan_scratch, t);
        bg_phase_zero_out(c, t, T);
        __syncthreads();
        bg_phase_emit(c, t, T);
        if (t == 0) {
            const uint32_t st = c.scal[BG_S_STATUS];
            a.out_len[b] = st ? 0u : 18u + c.scal[BG_S_PAYLOAD] + 8u;
            a.status[b] = st;
            if (st) atomicOr(a.err_flag, 1u);
        }
        fence_proxy_async();   /* our generic-proxy accesses to the payload area precede the next TMA write */
        __syncthreads();       /* everyone is done with the payload and the tables before the next block lands */
        PROF_MARK(9);
    }
}

/* ---- sizes -> exclusive offsets (single CTA, tiles of 1024 with a running carry).
 * count_dev (optional) overrides the element count, base_dev (optional) seeds the carry: both live on the
 * device so that batches chain without a host round trip. ---- */
__global__ void __launch_bounds__(1024, 1)
bgzf_scan_kernel(const uint32_t *len, uint64_t *off, uint32_t nmax, const uint64_t *count_dev, const uint64_t *base_dev,
                 uint64_t *total_out)
{
    __shared__ uint32_t v[1024];
    __shared__ uint32_t scratch[36];
    const uint32_t t = threadIdx.x;
    uint32_t n = nmax;
    if (count_dev) { const uint64_t cd = *count_dev; n = cd < nmax ? (uint32_t)cd : nmax; }
    uint64_t carry = base_dev ? *base_dev : 0ull;
    __syncthreads();   /* total_out may alias base_dev: everyone has read it before anyone writes */
    for (uint32_t start = 0; start < n; start += 1024) {
        const uint32_t i = start + t;
        v[t] = i < n ? len[i] : 0u;
        __syncthreads();
        const uint32_t total = block_exclusive_scan_1024(v, scratch, t);
        if (i < n) off[i] = carry + v[t];
        carry += total;
        __syncthreads();
    }
    if (t == 0) *total_out = carry;
}

/* ---- slots -> contiguous stream.  One CTA per member; destination may start at any byte. ---- */
__global__ void __launch_bounds__(256)
bgzf_gather_kernel(const uint8_t *slots, const uint32_t *len, const uint64_t *off, uint32_t nblocks, uint8_t *out,
                   const uint64_t *total, int append_eof)
{
    const uint32_t b = blockIdx.x, t = threadIdx.x;
    if (b == nblocks) {
        if (append_eof && t < 28) {
            const uint8_t eof[28] = { 0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                                      0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            out[*total + t] = eof[t];
        }
        return;
    }
    const uint32_t n = len[b];
    const uint32_t *src = (const uint32_t *)(slots + (size_t)b * BG_SLOT_BYTES);
    uint8_t *dst = out + off[b];
    const uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);   /* bytes until dst is word aligned */
    if (n <= head + 4) {
        for (uint32_t i = t; i < n; i += blockDim.x) dst[i] = ((const uint8_t *)src)[i];
        return;
    }
    if (t < head) dst[t] = ((const uint8_t *)src)[t];
    const uint32_t words = (n - head) >> 2;
    uint32_t *dw = (uint32_t *)(dst + head);
    const uint32_t sh = head * 8u;
    for (uint32_t i = t; i < words; i += blockDim.x)
        dw[i] = __funnelshift_r(src[i], src[i + 1], sh);       /* src word i+1 stays inside the 64 KiB slot + pad */
    const uint32_t done = head + words * 4u;
    if (t < n - done) dst[done + t] = ((const uint8_t *)src)[done + t];
}

extern "C" cudaError_t bgzf_launch_compress(const BgzfCompressArgs *a, int grid, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(bgzf_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
    if (e != cudaSuccess) return e;
    bgzf_compress_kernel<<<grid, BG_THREADS, SM_TOTAL, stream>>>(*a);
    return cudaGetLastError();
}

extern "C" cudaError_t bgzf_launch_scan(const uint32_t *len, uint64_t *off, uint32_t nmax, const uint64_t *count_dev,
                                        const uint64_t *base_dev, uint64_t *total_out, cudaStream_t stream)
{
    bgzf_scan_kernel<<<1, 1024, 0, stream>>>(len, off, nmax, count_dev, base_dev, total_out);
    return cudaGetLastError();
}

/* total (device) is read as the running base and updated to the new end of the stream */
extern "C" cudaError_t bgzf_launch_compact(const uint8_t *slots, const uint32_t *len, uint64_t *off, uint32_t nblocks,
                                           uint8_t *out, uint64_t *total, int append_eof, cudaStream_t stream)
{
    bgzf_scan_kernel<<<1, 1024, 0, stream>>>(len, off, nblocks, nullptr, total, total);
    bgzf_gather_kernel<<<nblocks + 1, 256, 0, stream>>>(slots, len, off, nblocks, out, total, append_eof);
    return cudaGetLastError();
}
