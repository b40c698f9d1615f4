/*
 * multi.c — one host buffer over several GPUs (the reference's fan-out: applet/7bgzf.c:159-227, a thread per
 * block).  BGZF blocks are independent, so GPU g of G takes the contiguous block range [B*g/G, B*(g+1)/G): no data-path
 * collective, no NCCL — every GPU runs its own pipelined host-buffer call (H2D, kernels, D2H) on its own context from its
 * own host thread, and the shard outputs are joined at host-known offsets.  The stream is byte-identical for any G
 * (per-block determinism), which the tests and bench.py check.
 *
 * Host code is plain C over the b200bgzf_* C ABI; no CUDA calls here.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200bgzf.h"

#define MULTI_MAX 64

struct b200bgzf_multi {
    int n;
    b200bgzf_ctx *ctx[MULTI_MAX];
};

int b200bgzf_multi_create(b200bgzf_multi **out, const int *devices, int ndevices)
{
    if (!out || ndevices < 1 || ndevices > MULTI_MAX) return B200BGZF_E_ARG;
    *out = NULL;
    b200bgzf_multi *m = (b200bgzf_multi *)calloc(1, sizeof *m);
    if (!m) return B200BGZF_E_ARG;
    for (int i = 0; i < ndevices; i++) {
        const int r = b200bgzf_create(&m->ctx[i], devices ? devices[i] : i);
        if (r != 0) {
            b200bgzf_multi_destroy(m);
            return r;
        }
        m->n = i + 1;
    }
    *out = m;
    return B200BGZF_OK;
}

void b200bgzf_multi_destroy(b200bgzf_multi *m)
{
    if (!m) return;
    for (int i = 0; i < MULTI_MAX; i++)
        if (m->ctx[i]) b200bgzf_destroy(m->ctx[i]);
    free(m);
}

int b200bgzf_multi_count(const b200bgzf_multi *m) { return m ? m->n : 0; }
b200bgzf_ctx *b200bgzf_multi_ctx(b200bgzf_multi *m, int i) { return m && i >= 0 && i < m->n ? m->ctx[i] : NULL; }

void b200bgzf_shard_blocks(uint64_t nblocks, int shard, int nshards, uint64_t *first, uint64_t *last)
{
    /* (128-bit product: 64 GiB inputs have ~2^20 blocks, nothing overflows, but stay exact for any size) */
    *first = (uint64_t)(((unsigned __int128)nblocks * (unsigned)shard) / (unsigned)nshards);
    *last = (uint64_t)(((unsigned __int128)nblocks * (unsigned)(shard + 1)) / (unsigned)nshards);
}

size_t b200bgzf_multi_compress_bound(const b200bgzf_multi *m, size_t in_bytes, uint32_t block_size)
{
    const size_t b = b200bgzf_compress_bound(in_bytes, block_size);
    return b && m ? b + (size_t)(m->n - 1) * B200BGZF_EOF_BYTES : b;
}

struct shard_job {
    pthread_t th;
    b200bgzf_ctx *ctx;
    const unsigned char *in;
    unsigned char *out;
    size_t in_bytes, out_cap, out_bytes;
    uint32_t block_size;
    int level, rc, inflate;
    unsigned flags;
};

static void *shard_main(void *arg)
{
    struct shard_job *j = (struct shard_job *)arg;
    j->out_bytes = 0;
    if (j->inflate)
        j->rc = b200bgzf_inflate_host(j->ctx, j->in, j->in_bytes, j->out, j->out_cap, &j->out_bytes, j->flags);
    else
        j->rc = b200bgzf_compress_host(j->ctx, j->in, j->in_bytes, j->block_size, j->level, j->out, j->out_cap, &j->out_bytes, j->flags);
    return NULL;
}

/* runs the jobs, the first one on the calling thread; returns the worst status (errors before "does not fit") */
static int run_jobs(struct shard_job *jobs, int n)
{
    for (int i = 1; i < n; i++)
        if (pthread_create(&jobs[i].th, NULL, shard_main, &jobs[i]) != 0) {
            shard_main(&jobs[i]);
            jobs[i].th = 0;
        }
    shard_main(&jobs[0]);
    int worst = 0;
    for (int i = 0; i < n; i++) {
        if (i && jobs[i].th) pthread_join(jobs[i].th, NULL);
        if (jobs[i].rc < 0 && worst >= 0) worst = jobs[i].rc;
        else if (jobs[i].rc > 0 && worst == 0) worst = jobs[i].rc;
    }
    return worst;
}

int b200bgzf_multi_compress_host(b200bgzf_multi *m, const void *in, size_t in_bytes, uint32_t block_size, int level, void *out,
                                 size_t out_cap, size_t *out_bytes, unsigned flags)
{
    if (!m || !out || !out_bytes || (!in && in_bytes) || block_size == 0 || block_size > B200BGZF_MAX_BLOCK_SIZE) return B200BGZF_E_ARG;
    if (out_cap < b200bgzf_multi_compress_bound(m, in_bytes, block_size)) return B200BGZF_E_NOSPACE;
    const uint64_t nb = (in_bytes + block_size - 1) / block_size;
    int n = m->n;
    if ((uint64_t)n > nb) n = nb ? (int)nb : 1;
    struct shard_job jobs[MULTI_MAX];
    memset(jobs, 0, sizeof jobs);
    for (int g = 0; g < n; g++) {
        uint64_t b0, b1;
        b200bgzf_shard_blocks(nb, g, n, &b0, &b1);
        const size_t byte0 = (size_t)b0 * block_size, byte1 = (size_t)b1 * block_size < in_bytes ? (size_t)b1 * block_size : in_bytes;
        /* where this shard may write: behind the worst case of everything before it */
        const size_t off = byte0 + (size_t)b0 * 38u + (size_t)g * B200BGZF_EOF_BYTES;
        jobs[g].ctx = m->ctx[g];
        jobs[g].in = (const unsigned char *)in + byte0;
        jobs[g].in_bytes = byte1 - byte0;
        jobs[g].out = (unsigned char *)out + off;
        jobs[g].out_cap = (byte1 - byte0) + (size_t)(b1 - b0) * 38u + B200BGZF_EOF_BYTES;
        jobs[g].block_size = block_size;
        jobs[g].level = level;
        jobs[g].flags = flags & ~B200BGZF_APPEND_EOF;
    }
    const int rc = run_jobs(jobs, n);
    if (rc < 0) return rc;
    /* join the shards at host-known offsets (shard 0 is in place) */
    size_t pos = jobs[0].out_bytes;
    for (int g = 1; g < n; g++) {
        memmove((unsigned char *)out + pos, jobs[g].out, jobs[g].out_bytes);
        pos += jobs[g].out_bytes;
    }
    if ((flags & B200BGZF_APPEND_EOF) && !(flags & B200BGZF_FRAME_MIGZ)) {
        static const unsigned char eof[28] = { 0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                                               0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        memcpy((unsigned char *)out + pos, eof, sizeof eof);
        pos += sizeof eof;
    }
    *out_bytes = pos;
    return rc;
}

int b200bgzf_multi_inflate_host(b200bgzf_multi *m, const void *in, size_t in_bytes, void *out, size_t out_cap, size_t *out_bytes,
                                unsigned flags)
{
    if (!m || !in || !out_bytes) return B200BGZF_E_ARG;
    /* the reference's header walk (applet/7bgzf.c:306-330) finds the members; shards are contiguous member ranges of
     * about equal compressed size, their outputs land at the running ISIZE sums */
    const unsigned char *p = (const unsigned char *)in;
    int n = m->n;
    size_t cut_in[MULTI_MAX + 1], cut_out[MULTI_MAX + 1];
    size_t off = 0, total = 0;
    int g = 0;
    cut_in[0] = 0;
    cut_out[0] = 0;
    while (off < in_bytes) {
        uint64_t sz = 0;
        if (!b200bgzf_member_header(p + off, in_bytes - off, &sz)) return B200BGZF_E_FORMAT;
        const unsigned char *t = p + off + sz - 4;
        total += (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
        off += (size_t)sz;
        while (g + 1 < n && off >= (size_t)(((unsigned __int128)in_bytes * (unsigned)(g + 1)) / (unsigned)n)) {
            g++;
            cut_in[g] = off;
            cut_out[g] = total;
        }
    }
    while (g + 1 <= n) {
        g++;
        cut_in[g] = off;
        cut_out[g] = total;
    }
    *out_bytes = total;
    if (total > out_cap || (!out && total)) return B200BGZF_E_NOSPACE;
    struct shard_job jobs[MULTI_MAX];
    memset(jobs, 0, sizeof jobs);
    int used = 0;
    for (int k = 0; k < n; k++) {
        if (cut_in[k + 1] == cut_in[k]) continue;
        jobs[used].ctx = m->ctx[used];
        jobs[used].in = p + cut_in[k];
        jobs[used].in_bytes = cut_in[k + 1] - cut_in[k];
        jobs[used].out = (unsigned char *)out + cut_out[k];
        jobs[used].out_cap = cut_out[k + 1] - cut_out[k];
        jobs[used].inflate = 1;
        jobs[used].flags = flags;
        used++;
    }
    return used ? run_jobs(jobs, used) : B200BGZF_OK;
}

/* ---- the other containers over several GPUs (SURVEY 8e for the rows of 8f): pieces are independent, GPU g takes a contiguous
 * range of them — whole members where the container has several — and the framing is written once over the joined stream ---- */
struct piece_job {
    pthread_t th;
    b200bgzf_ctx *ctx;
    const unsigned char *in;
    unsigned char *out;
    size_t in_bytes, cap, bytes, np;
    uint32_t bs;
    int level, rc;
    b200bgzf_piece_spec sp;
    uint64_t *off;
    uint32_t *crc;
};

static void *piece_main(void *arg)
{
    struct piece_job *j = (struct piece_job *)arg;
    j->bytes = 0;
    j->rc = b200bgzf_compress_pieces_host(j->ctx, j->in, j->in_bytes, j->bs, j->level, &j->sp, j->out, j->cap, &j->bytes, j->off, j->crc, j->np);
    return NULL;
}

size_t b200bgzf_multi_container_bound(const b200bgzf_multi *m, int kind, uint32_t param, size_t in_bytes)
{
    const size_t b = b200bgzf_container_bound(kind, param, in_bytes);
    return b && m ? b + (size_t)m->n * (B200BGZF_EOF_BYTES + 2u * B200BGZF_MAX_GAP) : b;
}

int b200bgzf_multi_container_compress_host(b200bgzf_multi *m, int kind, uint32_t param, const void *in, size_t in_bytes, int level,
                                           void *out, size_t out_cap, size_t *out_bytes)
{
    uint32_t bs;
    b200bgzf_piece_spec sp;
    if (!m || !out || !out_bytes || (!in && in_bytes)) return B200BGZF_E_ARG;
    if (b200bgzf_container_plan(kind, param, &bs, &sp) != 0) return B200BGZF_E_ARG;
    if (kind == B200BGZF_CONTAINER_RAZF && (in_bytes >> 32)) return B200BGZF_E_ARG;
    if (out_cap < b200bgzf_multi_container_bound(m, kind, param, in_bytes)) return B200BGZF_E_NOSPACE;
    *out_bytes = 0;
    const size_t np_total = (in_bytes + bs - 1) / bs;
    const size_t per_call = kind == B200BGZF_CONTAINER_DICTZIP ? 32762u : (np_total ? np_total : 1);   /* dictzip: one member per round */
    const size_t arr = per_call < np_total ? per_call : np_total;
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * arr + 8);
    uint32_t *crc = (uint32_t *)malloc(sizeof(uint32_t) * arr + 8);
    if (!off || !crc) { free(off); free(crc); return B200BGZF_E_ARG; }
    int rc = B200BGZF_OK;
    size_t pos = 0, done = 0;
    do {
        const size_t np = np_total - done < per_call ? np_total - done : per_call;
        const size_t byte0 = done * bs, bytes = (done + np) * (size_t)bs < in_bytes ? np * (size_t)bs : in_bytes - byte0;
        const size_t head = b200bgzf_container_head(kind, np);
        size_t stream = 0;
        if (np) {
            /* what is dealt out: single pieces of a one-member container, whole members otherwise */
            const uint64_t k = sp.member_blocks == 0xffffffffu ? 1u : sp.member_blocks;
            const uint64_t groups = (np + k - 1) / k;
            int n = m->n;
            if ((uint64_t)n > groups) n = (int)groups;
            struct piece_job jobs[MULTI_MAX];
            memset(jobs, 0, sizeof jobs);
            size_t place = pos + head;
            for (int g = 0; g < n; g++) {
                uint64_t g0, g1;
                b200bgzf_shard_blocks(groups, g, n, &g0, &g1);
                const size_t p0 = (size_t)(g0 * k), p1 = (size_t)(g1 * k) < np ? (size_t)(g1 * k) : np;
                struct piece_job *j = &jobs[g];
                j->ctx = m->ctx[g];
                j->bs = bs;
                j->level = level;
                j->sp = sp;
                j->sp.piece_base = p0;
                j->sp.piece_total = np;
                j->in = (const unsigned char *)in + byte0 + p0 * (size_t)bs;
                j->in_bytes = (p1 * (size_t)bs < bytes ? p1 * (size_t)bs : bytes) - p0 * (size_t)bs;
                j->np = p1 - p0;
                j->off = off + p0;
                j->crc = crc + p0;
                j->out = (unsigned char *)out + place;       /* behind the worst case of the shards before it */
                j->cap = b200bgzf_compress_bound(j->in_bytes, bs) + b200bgzf_pieces_gap_bytes(j->in_bytes, bs, &j->sp);
                place += j->cap;
            }
            if (place > out_cap) { rc = B200BGZF_E_NOSPACE; break; }
            for (int g = 1; g < n; g++)
                if (pthread_create(&jobs[g].th, NULL, piece_main, &jobs[g]) != 0) {
                    piece_main(&jobs[g]);
                    jobs[g].th = 0;
                }
            piece_main(&jobs[0]);
            for (int g = 0; g < n; g++) {
                if (g && jobs[g].th) pthread_join(jobs[g].th, NULL);
                if (jobs[g].rc < 0 && rc >= 0) rc = jobs[g].rc;
                else if (jobs[g].rc > 0 && rc == 0) rc = jobs[g].rc;
            }
            if (rc != 0) break;
            /* join the shard streams (shard 0 is in place); piece offsets become offsets in the joined stream */
            for (int g = 0; g < n; g++) {
                if (g) memmove((unsigned char *)out + pos + head + stream, jobs[g].out, jobs[g].bytes);
                for (size_t i = 0; i < jobs[g].np; i++) jobs[g].off[i] += stream;
                stream += jobs[g].bytes;
            }
        }
        const size_t total = b200bgzf_container_frame(kind, param, (unsigned char *)out + pos, out_cap - pos, stream, off, crc, np, bytes);
        if (!total && !(kind == B200BGZF_CONTAINER_MIGZ && np == 0)) { rc = B200BGZF_E_NOSPACE; break; }
        pos += total;
        done += np;
    } while (done < np_total);
    free(off);
    free(crc);
    if (rc == B200BGZF_E_NOFIT && kind == B200BGZF_CONTAINER_MIGZ && !(param & B200BGZF_PARAM_SAFE))
        return b200bgzf_multi_container_compress_host(m, kind, param | B200BGZF_PARAM_SAFE, in, in_bytes, level, out, out_cap, out_bytes);
    if (rc == 0) *out_bytes = pos;
    return rc;
}
