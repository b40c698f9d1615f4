/*
 * hook_mt.c — stand-in for htslib's thread pool: N pthreads, each calling bgzf_compress() from a given shared
 * object (the GPU 7bgzf.so, or the reference's) on one <= 0xff00-byte block at a time, over a file or a
 * synthetic buffer.  Measurement tool only ("emulate htslib with an N-thread harness").
 *
 *   hook_mt <7bgzf.so> <threads> <input file> [repeat]
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int (*bgzf_compress_fn)(void *, size_t *, const void *, size_t, int);
static bgzf_compress_fn fn;
static const uint8_t *data;
static size_t nbytes, nblocks;
static int nthreads, repeat = 1;
static uint64_t out_total[256];

static void *worker(void *arg)
{
    const int tid = (int)(intptr_t)arg;
    uint8_t dst[65536];
    uint64_t total = 0;
    for (int r = 0; r < repeat; r++)
        for (size_t b = (size_t)tid; b < nblocks; b += (size_t)nthreads) {
            const size_t off = b * 0xff00u, len = nbytes - off < 0xff00u ? nbytes - off : 0xff00u;
            size_t dl = sizeof dst;
            if (fn(dst, &dl, data + off, len, 6) != 0) { fprintf(stderr, "bgzf_compress failed at block %zu\n", b); exit(1); }
            total += dl;
        }
    out_total[tid] = total;
    return NULL;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: hook_mt <so> <threads> <file> [repeat]\n"); return 2; }
    void *dl = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!dl) { fprintf(stderr, "%s\n", dlerror()); return 1; }
    fn = (bgzf_compress_fn)dlsym(dl, "bgzf_compress");
    if (!fn) { fprintf(stderr, "no bgzf_compress in %s\n", argv[1]); return 1; }
    nthreads = atoi(argv[2]);
    if (nthreads < 1 || nthreads > 256) return 2;
    if (argc > 4) repeat = atoi(argv[4]);
    FILE *f = fopen(argv[3], "rb");
    if (!f) { perror(argv[3]); return 1; }
    fseek(f, 0, SEEK_END);
    nbytes = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = malloc(nbytes ? nbytes : 1);
    if (fread(buf, 1, nbytes, f) != nbytes) return 1;
    fclose(f);
    data = buf;
    nblocks = (nbytes + 0xff00u - 1) / 0xff00u;
    { uint8_t dst[65536]; size_t dlen = sizeof dst; fn(dst, &dlen, data, nbytes < 0xff00u ? nbytes : 0xff00u, 6); }   /* initialisation outside the timing */
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t th[256];
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, (void *)(intptr_t)i);
    uint64_t total = 0;
    for (int i = 0; i < nthreads; i++) { pthread_join(th[i], NULL); total += out_total[i]; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
    printf("{\"threads\": %d, \"MB_per_s\": %.1f, \"us_per_call\": %.0f, \"ratio\": %.4f, \"bytes\": %zu}\n", nthreads,
           (double)nbytes * repeat / dt / 1e6, dt / ((double)nblocks * repeat / nthreads) * 1e6, (double)total / ((double)nbytes * repeat), nbytes);
    return 0;
}
