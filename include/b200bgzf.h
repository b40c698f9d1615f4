/*
 * b200bgzf.h — C ABI of the B200-native BGZF codec (lib7bgzf_b200.so / 7bgzf.so).
 *
 * This is the drop-in boundary for the one hot path of cielavenir/7bgzf: per-64 KiB-block BGZF compress
 * (libdeflate level class 1..12) and BGZF inflate.  Plain pointers and sizes only; every entry point states
 * the reference interface it stands in for (paths relative to the reference tree).
 *
 * There is NO CPU fallback: every compress/inflate call runs the sm_100a kernels and fails with
 * B200BGZF_E_CUDA when no usable GPU (or the CUDA runtime) is present.
 */
#ifndef B200BGZF_H
#define B200BGZF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200BGZF_BLOCK_SIZE 0xff00u      /* htslib BGZF_BLOCK_SIZE; applet/7bgzf.c:146-147 */
#define B200BGZF_MAX_BLOCK_SIZE 0x10000u /* htslib BGZF_MAX_BLOCK_SIZE = largest member and largest payload */
#define B200BGZF_EOF_BYTES 28u           /* bgzf_compress.c:43-49 */

/* return codes.  0 and 1 and -1 keep the meaning they have in bgzf_compress.c:39-198 */
#define B200BGZF_OK 0
#define B200BGZF_E_NOFIT 1      /* a member would not fit its capacity (reference: "libdeflate_deflate 1", returns 1) */
#define B200BGZF_E_ARG (-1)     /* bad argument / capacity below the fixed minimum (reference returns -1) */
#define B200BGZF_E_CUDA (-2)    /* no device, out of device memory, launch failure */
#define B200BGZF_E_FORMAT (-3)  /* input is not BGZF / corrupt DEFLATE data */
#define B200BGZF_E_NOSPACE (-4) /* output buffer too small */
#define B200BGZF_E_CRC (-5)     /* CRC32 or ISIZE mismatch (only with B200BGZF_VERIFY) */

#define B200BGZF_APPEND_EOF 1u  /* compress: finish the stream with the 28-byte EOF member (applet/7bgzf.c:283-289) */
#define B200BGZF_VERIFY 2u      /* inflate: check CRC32 and ISIZE of every member (the reference does not: 7bgzf.c:350-354) */
#define B200BGZF_FRAME_MIGZ 4u  /* compress: MiGz members instead of BGZF ones — the same DEFLATE data behind the gzip subfield "MZ" with the
                                   compressed size as u32 (applet/7migz.c:133-243).  No EOF marker exists in that
                                   container (B200BGZF_APPEND_EOF is ignored); block_size <= 64512 keeps every member inside its slot. */

typedef struct b200bgzf_ctx b200bgzf_ctx;

/* One context = one GPU, its streams, pinned staging and device workspaces.  device < 0: current device.
 * Replaces the per-call libdeflate_alloc_compressor/free of lib/zlibutil.c:186-188 with process-lifetime pools. */
int b200bgzf_create(b200bgzf_ctx **ctx, int device);
void b200bgzf_destroy(b200bgzf_ctx *ctx);
const char *b200bgzf_strerror(int code);
/* last CUDA error text seen by this context (empty string if none) */
const char *b200bgzf_last_error(const b200bgzf_ctx *ctx);

/* Worst-case size of the BGZF stream for in_bytes of payload cut into block_size blocks (+EOF). */
size_t b200bgzf_compress_bound(size_t in_bytes, uint32_t block_size);

/*
 * Compress a contiguous buffer into a contiguous BGZF stream; block b carries payload bytes
 * [b*block_size, min((b+1)*block_size, in_bytes)).  This is the applet's _compress() loop
 * (applet/7bgzf.c:133-293: read block, libdeflate_deflate, frame, write in order) as one batched call.
 *   *_device: d_in/d_out are device pointers (d_in 16-byte aligned for the TMA path), `stream` is a
 *             cudaStream_t passed as void* (NULL: the context's stream).  Synchronises before returning.
 *   *_host:   host pointers (pinned memory copies fastest); pipelined H2D / kernels / D2H.
 * level: 1..12 (libdeflate classes, lib/libdeflate/deflate_compress.c:3921-4007).
 */
int b200bgzf_compress_device(b200bgzf_ctx *ctx, const void *d_in, size_t in_bytes, uint32_t block_size, int level,
                             void *d_out, size_t out_cap, size_t *out_bytes, unsigned flags, void *stream);
int b200bgzf_compress_host(b200bgzf_ctx *ctx, const void *in, size_t in_bytes, uint32_t block_size, int level, void *out,
                           size_t out_cap, size_t *out_bytes, unsigned flags);

/*
 * SURVEY §8(f) rank 2 — the index side-product.  Same as b200bgzf_compress_host, and also reports where every member
 * starts: member_off[b] = byte offset in `out` of the member that carries block b (member_cap >= number of blocks;
 * the EOF marker is not listed).  These are the offsets the device scan computes for the compaction anyway, copied
 * back with each batch.  Together with b*block_size they are the (compressed, uncompressed) address pairs of
 * htslib's .gzi index and the upper 48 bits of BAM virtual offsets.  The reference has no counterpart (its applet
 * writes no index: applet/7bgzf.c:133-293); the .gzi layout follows htslib's bgzf_index_dump (not in the reference
 * tree, so parity with htslib is unpinned — tests check the index against a header walk of the stream).
 */
int b200bgzf_compress_host_index(b200bgzf_ctx *ctx, const void *in, size_t in_bytes, uint32_t block_size, int level, void *out,
                                 size_t out_cap, size_t *out_bytes, unsigned flags, uint64_t *member_off, size_t member_cap);
/* Serialise a .gzi: u64 entry count, then (caddr, uaddr) for every member except the first (which is 0,0 by
 * definition), little endian.  Needs 8 + 16*(nmembers-1) bytes; returns the bytes written or 0 if cap is too small. */
size_t b200bgzf_gzi_format(const uint64_t *caddr, const uint64_t *uaddr, size_t nmembers, void *dst, size_t cap);

/*
 * Compress nblocks independent payloads (src[i], slen[i] <= 65536) into dst[i]; dlen[i] holds the capacity on
 * entry and the member size on return; status[i] gets the per-block code of bgzf_compress().  This is the
 * batch form of bgzf_compress.c:39-198.  nblocks == 1 is what the LD_PRELOAD hook calls for every htslib block:
 * that case takes a dedicated low-latency path (one copy in, one kernel, one copy out, one synchronisation) on
 * one of 128 independent lanes, so concurrent callers each drive their own stream; while few callers are in
 * flight the kernel of a call runs on a thread-block cluster of 8 (or 4) SMs that share the block's match search
 * (same output bytes, less than half the latency); callers beyond the host's core count wait by sleeping, not
 * spinning.
 */
int b200bgzf_compress_blocks_host(b200bgzf_ctx *ctx, const void *const *src, const uint32_t *slen, void *const *dst,
                                  size_t *dlen, int *status, uint32_t nblocks, int level);

/*
 * Inflate a BGZF stream (any sequence of BGZF members, EOF markers included) — the applet's _decompress()
 * (applet/7bgzf.c:295-365) with zlibutil_auto_inflate (lib/zlibutil.c:82-93) as the per-member decoder.
 * The *_size helper walks the member headers on the host (7bgzf.c:81-131) and returns the member count and the
 * total ISIZE so the caller can size the output.
 */
int b200bgzf_inflate_size_host(const void *in, size_t in_bytes, size_t *out_bytes, size_t *nmembers);
/* The member header parser of that loop (_read_gz_header, applet/7bgzf.c:81-131), every flavour it accepts: BGZF ("BC"),
 * MiGz ("MZ"), mgzip v1/v2 ("IG"), jerodsanto's mgzip; optional name / comment / header-CRC fields are skipped.  Returns
 * the header length (offset of the DEFLATE data) and stores the whole member size, or returns 0 if `p` does not start
 * such a member (or it is cut short by `avail`).  The *_host inflate entry points accept all of these; the
 * device-resident one, which finds members by their signature on the device, takes BGZF only. */
uint32_t b200bgzf_member_header(const void *p, size_t avail, uint64_t *member_bytes);
int b200bgzf_inflate_device(b200bgzf_ctx *ctx, const void *d_in, size_t in_bytes, void *d_out, size_t out_cap,
                            size_t *out_bytes, unsigned flags, void *stream);
int b200bgzf_inflate_host(b200bgzf_ctx *ctx, const void *in, size_t in_bytes, void *out, size_t out_cap, size_t *out_bytes,
                          unsigned flags);

/*
 * Several GPUs behind one call (the reference's fan-out over blocks: applet/7bgzf.c:159-227).  GPU g of G takes
 * the contiguous block range [B*g/G, B*(g+1)/G) of the input (b200bgzf_shard_blocks), runs its own pipelined host-buffer
 * call on its own context and host thread, and the shard outputs are joined at host-known offsets: no collective, and the
 * stream is byte-identical for every G.  devices == NULL: ordinals 0..ndevices-1 (an ordinal may repeat: two contexts on
 * one GPU).  The output buffer must hold b200bgzf_multi_compress_bound() bytes (28 more per extra shard than the
 * single-GPU bound: every shard is placed behind the worst case of the ones before it, then moved down).
 */
typedef struct b200bgzf_multi b200bgzf_multi;
int b200bgzf_multi_create(b200bgzf_multi **m, const int *devices, int ndevices);
void b200bgzf_multi_destroy(b200bgzf_multi *m);
int b200bgzf_multi_count(const b200bgzf_multi *m);
b200bgzf_ctx *b200bgzf_multi_ctx(b200bgzf_multi *m, int i);
void b200bgzf_shard_blocks(uint64_t nblocks, int shard, int nshards, uint64_t *first, uint64_t *last);
size_t b200bgzf_multi_compress_bound(const b200bgzf_multi *m, size_t in_bytes, uint32_t block_size);
int b200bgzf_multi_compress_host(b200bgzf_multi *m, const void *in, size_t in_bytes, uint32_t block_size, int level, void *out,
                                 size_t out_cap, size_t *out_bytes, unsigned flags);
int b200bgzf_multi_inflate_host(b200bgzf_multi *m, const void *in, size_t in_bytes, void *out, size_t out_cap, size_t *out_bytes,
                                unsigned flags);

/* Page-locked host memory for the *_host entry points (pageable buffers work too, but are copied at a fraction of
 * the PCIe rate).  The applet reads stdin straight into such buffers. */
void *b200bgzf_host_alloc(size_t bytes);
void b200bgzf_host_free(void *p);

/* Per-phase cycle counters of the compress kernel (development aid; n <= 16). Resets them when reset != 0. */
int b200bgzf_profile(b200bgzf_ctx *ctx, int enable, unsigned long long *cycles, int n, int reset);
/* number of kernel launches issued by this context so far */
unsigned long long b200bgzf_launch_count(const b200bgzf_ctx *ctx);

/*
 * The htslib hook (exported from 7bgzf.so only).  Same signature and return conventions as the reference's
 * bgzf_compress.c:39: 0 ok; -1 if *dlen < 28 for slen == 0 or *dlen < 26; 1 (+ "libdeflate_deflate 1" on
 * stderr) when the member does not fit; `level` is ignored, BGZF_METHOD=<name><digits> selects the level.
 */
int bgzf_compress(void *dst, size_t *dlen, const void *src, size_t slen, int level);

/* BGZF_METHOD parser used by the hook (bgzf_compress.c:53-113).  Returns 0 and the level (1..12) to use,
 * or -1 if the digits are out of range.  Every method name maps onto this codec; unset/zlib -> 6. */
int b200bgzf_parse_method(const char *spec, int *level, char *method_name, size_t method_cap);

#ifdef __cplusplus
}
#endif
#endif
