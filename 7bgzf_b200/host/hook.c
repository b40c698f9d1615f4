/*
 * hook.c — the LD_PRELOAD entry point: htslib's
 *     int bgzf_compress(void *dst, size_t *dlen, const void *src, size_t slen, int level)
 * with the return conventions of the reference's bgzf_compress.c:39-198, served by the GPU codec.
 *
 *   BGZF_METHOD=libdeflate6 LD_PRELOAD=./7bgzf.so samtools view -b ...      (samtools linked to libhts.so)
 *
 * Host code is plain C; it reaches CUDA only through the b200bgzf_* C ABI.  Concurrent callers (htslib's
 * thread pool) are safe: initialisation is a pthread_once, and each call borrows its own stream/SM lane.
 * If no GPU can be initialised every non-empty call fails with -1 and a line on stderr: no CPU fallback.
 * Optional: B200BGZF_DEVICE=<ordinal>.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200bgzf.h"

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static b200bgzf_ctx *g_ctx;
static int g_level = -1;
static int g_init_rc = B200BGZF_E_CUDA;

static const unsigned char k_eof[28] = { 0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                         0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0 };

static void hook_init(void)
{
    const char *dev = getenv("B200BGZF_DEVICE");
    if (b200bgzf_parse_method(getenv("BGZF_METHOD"), &g_level, NULL, 0) != 0) {
        fprintf(stderr, "b200bgzf: BGZF_METHOD level out of range (1..12)\n");
        g_level = -1;
        g_init_rc = B200BGZF_E_ARG;
        return;
    }
    g_init_rc = b200bgzf_create(&g_ctx, dev && *dev ? atoi(dev) : -1);
    if (g_init_rc != 0)
        fprintf(stderr, "b200bgzf: cannot initialise the GPU codec: %s\n", b200bgzf_strerror(g_init_rc));
}

__attribute__((destructor)) static void hook_fini(void)
{
    if (g_ctx) b200bgzf_destroy(g_ctx);
    g_ctx = NULL;
}

int bgzf_compress(void *dst, size_t *dlen, const void *src, size_t slen, int level_unused)
{
    (void)level_unused;
    if (!slen) {
        if (*dlen < 28) return -1;
        memcpy(dst, k_eof, 28);
        *dlen = 28;
        return 0;
    }
    pthread_once(&g_once, hook_init);
    if (*dlen < 26) return -1;
    if (g_init_rc != 0 || !g_ctx) return -1;
    if (slen > B200BGZF_MAX_BLOCK_SIZE) {
        fprintf(stderr, "libdeflate_deflate %d\n", 1);
        return 1;
    }
    const void *srcs[1] = { src };
    void *dsts[1] = { dst };
    uint32_t slens[1] = { (uint32_t)slen };
    size_t caps[1] = { *dlen > B200BGZF_MAX_BLOCK_SIZE ? B200BGZF_MAX_BLOCK_SIZE : *dlen };
    int st[1] = { 0 };
    int r = b200bgzf_compress_blocks_host(g_ctx, srcs, slens, dsts, caps, st, 1, g_level);
    if (r == B200BGZF_E_NOFIT) {
        fprintf(stderr, "libdeflate_deflate %d\n", 1);
        return 1;
    }
    if (r != 0) {
        fprintf(stderr, "b200bgzf: %s (%s)\n", b200bgzf_strerror(r), b200bgzf_last_error(g_ctx));
        return -1;
    }
    *dlen = caps[0];
    return 0;
}
