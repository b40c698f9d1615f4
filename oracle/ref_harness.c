/*
 * ref_harness.c — drives the UNMODIFIED reference built by oracle/Makefile.ref (oracle/_ref/7bgzf_ref.so).
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never linked into the product.
 *
 *   refh_open(so, level)       dlopen a private copy of the reference .so with BGZF_METHOD=libdeflate<level> latched
 *                              (the reference caches method/level in statics on first use, bgzf_compress.c:36-37,53)
 *   refh_compress(...)         N pthreads, each calling the reference's bgzf_compress() on a contiguous range of
 *                              0xff00-byte blocks — exactly what htslib's thread pool does ("fair" baseline)
 *   refh_inflate(...)          N pthreads calling the reference's libdeflate_deflate_decompress per member
 *   refh_crc32(...)            the reference's zlib crc32
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef int (*bgzf_compress_fn)(void *, size_t *, const void *, size_t, int);
typedef void *(*alloc_dec_fn)(void);
typedef int (*dec_fn)(void *, const void *, size_t, void *, size_t, size_t *);
typedef void (*free_dec_fn)(void *);
typedef unsigned long (*crc32_fn)(unsigned long, const unsigned char *, unsigned);

typedef struct {
    void *dl;
    bgzf_compress_fn compress;
    alloc_dec_fn alloc_dec;
    dec_fn dec;
    free_dec_fn free_dec;
    crc32_fn crc32;
    int level;
} refh;

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

refh *refh_open(const char *so_path, int level)
{
    /* private copy => private statics, so several levels can coexist in one process */
    char tmp[256], env[64];
    snprintf(tmp, sizeof tmp, "/tmp/refh_%d_L%d_%ld.so", (int)getpid(), level, (long)random());
    FILE *fi = fopen(so_path, "rb"), *fo = fi ? fopen(tmp, "wb") : NULL;
    if (!fi || !fo) { if (fi) fclose(fi); return NULL; }
    char buf[65536];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, fi)) > 0) fwrite(buf, 1, r, fo);
    fclose(fi);
    fclose(fo);
    refh *h = calloc(1, sizeof *h);
    h->dl = dlopen(tmp, RTLD_NOW | RTLD_LOCAL);
    unlink(tmp);
    if (!h->dl) { fprintf(stderr, "refh_open: %s\n", dlerror()); free(h); return NULL; }
    h->compress = (bgzf_compress_fn)dlsym(h->dl, "bgzf_compress");
    h->alloc_dec = (alloc_dec_fn)dlsym(h->dl, "libdeflate_alloc_decompressor");
    h->dec = (dec_fn)dlsym(h->dl, "libdeflate_deflate_decompress");
    h->free_dec = (free_dec_fn)dlsym(h->dl, "libdeflate_free_decompressor");
    h->crc32 = (crc32_fn)dlsym(h->dl, "crc32");
    h->level = level;
    if (!h->compress || !h->dec) { dlclose(h->dl); free(h); return NULL; }
    /* latch the level */
    snprintf(env, sizeof env, "libdeflate%d", level);
    const char *old = getenv("BGZF_METHOD");
    char *saved = old ? strdup(old) : NULL;
    setenv("BGZF_METHOD", env, 1);
    uint8_t dst[256];
    size_t dl = sizeof dst;
    h->compress(dst, &dl, "latch-the-method", 16, 0);
    if (saved) { setenv("BGZF_METHOD", saved, 1); free(saved); }
    else unsetenv("BGZF_METHOD");
    return h;
}

void refh_close(refh *h)
{
    if (!h) return;
    dlclose(h->dl);
    free(h);
}

/* single call passthrough (edge-case tests): returns the reference's return code */
int refh_bgzf_compress(refh *h, void *dst, size_t *dlen, const void *src, size_t slen) { return h->compress(dst, dlen, src, slen, 0); }

uint32_t refh_crc32(refh *h, const uint8_t *p, size_t n) { return (uint32_t)h->crc32(0, p, (unsigned)n); }

typedef struct {
    refh *h;
    const uint8_t *src;
    size_t n, block, b0, b1;
    uint8_t *slots;      /* 65536 per block, or NULL: discard output */
    uint32_t *sizes;
    int rc;
} cjob;

static void *cworker(void *arg)
{
    cjob *j = arg;
    uint8_t local[65536];
    for (size_t b = j->b0; b < j->b1; b++) {
        size_t off = b * j->block, len = j->n - off < j->block ? j->n - off : j->block, dl = 65536;
        uint8_t *dst = j->slots ? j->slots + b * 65536 : local;
        int r = j->h->compress(dst, &dl, j->src + off, len, 0);
        if (r) { j->rc = r; dl = 0; }
        j->sizes[b] = (uint32_t)dl;
    }
    return NULL;
}

/* returns seconds; sizes[b] = member size; slots may be NULL */
double refh_compress(refh *h, const uint8_t *src, size_t n, size_t block, int nthreads, uint8_t *slots, uint32_t *sizes, int *rc)
{
    size_t nb = (n + block - 1) / block;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > nb) nthreads = nb ? (int)nb : 1;
    pthread_t *th = calloc(nthreads, sizeof *th);
    cjob *jobs = calloc(nthreads, sizeof *jobs);
    double t0 = now();
    for (int i = 0; i < nthreads; i++) {
        jobs[i] = (cjob){ h, src, n, block, nb * i / nthreads, nb * (i + 1) / nthreads, slots, sizes, 0 };
        pthread_create(&th[i], NULL, cworker, &jobs[i]);
    }
    *rc = 0;
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], NULL);
        if (jobs[i].rc) *rc = jobs[i].rc;
    }
    double t = now() - t0;
    free(th);
    free(jobs);
    return t;
}

typedef struct {
    refh *h;
    const uint8_t *in;
    const uint64_t *in_off, *out_off;
    size_t m0, m1;
    uint8_t *out;
    int rc;
} djob;

static void *dworker(void *arg)
{
    djob *j = arg;
    void *d = j->h->alloc_dec();
    for (size_t m = j->m0; m < j->m1; m++) {
        const uint8_t *p = j->in + j->in_off[m];
        size_t msize = j->in_off[m + 1] - j->in_off[m], isize = j->out_off[m + 1] - j->out_off[m], got = 0;
        int r = j->h->dec(d, p + 18, msize - 26, j->out + j->out_off[m], isize, &got);
        if (r || got != isize) j->rc = r ? r : 100;
    }
    j->h->free_dec(d);
    return NULL;
}

/* inflates a BGZF stream with the reference's libdeflate decoder; returns seconds (<0: not BGZF) */
double refh_inflate(refh *h, const uint8_t *in, size_t n, int nthreads, uint8_t *out, size_t out_cap, size_t *out_len, int *rc)
{
    size_t cap = 1024, nm = 0;
    uint64_t *in_off = malloc(cap * 8), *out_off = malloc(cap * 8);
    size_t ip = 0, op = 0;
    while (ip < n) {
        if (n - ip < 28 || in[ip] != 0x1f || in[ip + 1] != 0x8b || in[ip + 12] != 'B' || in[ip + 13] != 'C') { free(in_off); free(out_off); return -1; }
        size_t msize = (size_t)(in[ip + 16] | (in[ip + 17] << 8)) + 1;
        if (ip + msize > n) { free(in_off); free(out_off); return -1; }
        const uint8_t *t = in + ip + msize - 4;
        size_t isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
        if (nm + 2 > cap) { cap *= 2; in_off = realloc(in_off, cap * 8); out_off = realloc(out_off, cap * 8); }
        in_off[nm] = ip;
        out_off[nm] = op;
        nm++;
        ip += msize;
        op += isize;
    }
    in_off[nm] = ip;
    out_off[nm] = op;
    *out_len = op;
    *rc = 0;
    if (op > out_cap) { *rc = 3; free(in_off); free(out_off); return 0; }
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > nm) nthreads = nm ? (int)nm : 1;
    pthread_t *th = calloc(nthreads, sizeof *th);
    djob *jobs = calloc(nthreads, sizeof *jobs);
    double t0 = now();
    for (int i = 0; i < nthreads; i++) {
        jobs[i] = (djob){ h, in, in_off, out_off, nm * i / nthreads, nm * (i + 1) / nthreads, out, 0 };
        pthread_create(&th[i], NULL, dworker, &jobs[i]);
    }
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], NULL);
        if (jobs[i].rc) *rc = jobs[i].rc;
    }
    double t = now() - t0;
    free(th);
    free(jobs);
    free(in_off);
    free(out_off);
    return t;
}
