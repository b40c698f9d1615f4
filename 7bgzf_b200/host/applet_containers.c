/*
 * applet_containers.c — the applet under its other names: 7gzip, 7gzinga, 7dictzip, 7razf, and 7migz with members of more
 * than 63 KiB (the reference is a multi-call binary too: cielbox.c:215-223).  Command lines follow the reference's:
 *   7gzip    -cl6 < in > out.gz          7gzip    -d < in.gz > out            (applet/7gzip.c, zlibrawstdio_compress.h)
 *   7migz    -cl6 [-b KiB] < in > out    7migz    -d < in > out               (applet/7migz.c:318-500)
 *   7gzinga  -cl6 < in > out.gzi         7gzinga  -cd in.gzi > out            (applet/7gzinga.c:310-495)
 *   7dictzip -cl6 [-X] in out.dz         7dictzip -cd in.dz > out             (applet/7dictzip.c:402-603; -X: 65280-byte chunks)
 *   7razf    -cl6 in > out.raz           7razf    -cd in.raz > out            (applet/7razf.c:387-580)
 * (a file operand is also accepted where the reference reads stdin, and the other way round).  These containers carry a
 * whole-file index or checksum, so the file is held in memory: read, one b200bgzf_container_* call, write.
 * Host code is C over the b200bgzf_* C ABI; no GPU => error, no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/b200bgzf.h"

static unsigned char *slurp(FILE *f, size_t *n)
{
    struct stat st;
    size_t cap = 1 << 20, len = 0;
    if (fstat(fileno(f), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) cap = (size_t)st.st_size + 1;
    unsigned char *buf = (unsigned char *)malloc(cap);
    while (buf) {
        const size_t got = fread(buf + len, 1, cap - len, f);
        len += got;
        if (len < cap) break;                       /* short read: end of file (or error) */
        unsigned char *nb = (unsigned char *)realloc(buf, cap * 2);
        if (!nb) { free(buf); buf = NULL; break; }
        buf = nb;
        cap *= 2;
    }
    *n = len;
    return buf;
}

/* kind: B200BGZF_CONTAINER_*; files: the operands left after the options */
int container_applet(const char *name, int kind, int decompress, int level, unsigned param, int ndevices, int nfiles, char **files)
{
    FILE *in = stdin, *out = stdout;
    if (nfiles > 0 && !(in = fopen(files[0], "rb"))) { fprintf(stderr, "failed to open %s\n", files[0]); return 2; }
    if (nfiles > 1 && !decompress && !(out = fopen(files[1], "wb"))) { fprintf(stderr, "failed to open %s\n", files[1]); return 2; }
    if (isatty(fileno(in)) || isatty(fileno(out))) { fprintf(stderr, "%s: refusing to read from / write to a terminal\n", name); return -1; }
    size_t n = 0;
    unsigned char *src = slurp(in, &n);
    if (!src) { fprintf(stderr, "out of memory\n"); return 1; }

    b200bgzf_ctx *ctx = NULL;
    b200bgzf_multi *multi = NULL;                       /* --devices=N: the pieces are dealt out over N GPUs (compress) */
    const char *dev = getenv("B200BGZF_DEVICE");
    int r;
    if (ndevices > 1) {
        r = b200bgzf_multi_create(&multi, NULL, ndevices);
        ctx = b200bgzf_multi_ctx(multi, 0);
    } else {
        r = b200bgzf_create(&ctx, dev && *dev ? atoi(dev) : -1);
    }
    if (r != 0) { fprintf(stderr, "b200bgzf: cannot initialise the GPU codec: %s\n", b200bgzf_strerror(r)); free(src); return 1; }

    unsigned char *dst = NULL;
    size_t cap = 0, produced = 0;
    long units = 0;
    int ret = 0;
    if (decompress) {
        size_t nu = 0;
        r = n ? b200bgzf_container_inflate_size(kind, src, n, &cap, &nu) : B200BGZF_E_FORMAT;
        units = (long)nu;
        if (r != 0) { fprintf(stderr, "%s: not a %s file (possibly corrupted)\n", name, name + 1); ret = 1; }
        else if (!(dst = (unsigned char *)malloc(cap + 1))) { fprintf(stderr, "out of memory\n"); ret = 1; }
        else {
            memset(dst, 0, cap + 1);
            r = b200bgzf_container_inflate_host(ctx, kind, src, n, dst, cap, &produced, B200BGZF_VERIFY);     /* unlike the reference, check every member's CRC-32 */
            if (r != 0) { fprintf(stderr, "inflate %d\n", r); ret = 1; }
        }
    } else {
        cap = multi ? b200bgzf_multi_container_bound(multi, kind, param, n) : b200bgzf_container_bound(kind, param, n);
        if (!cap || !(dst = (unsigned char *)malloc(cap))) { fprintf(stderr, cap ? "out of memory\n" : "%s: bad block size\n", name); ret = 1; }
        else {
            memset(dst, 0, cap);                 /* fault the pages in before the device copies into them */
            r = multi ? b200bgzf_multi_container_compress_host(multi, kind, param, src, n, level, dst, cap, &produced)
                      : b200bgzf_container_compress_host(ctx, kind, param, src, n, level, dst, cap, &produced);
            if (r == B200BGZF_E_NOFIT) { fprintf(stderr, "libdeflate_deflate %d\n", 1); ret = 1; }
            else if (r != 0) { fprintf(stderr, "b200bgzf: %s (%s)\n", b200bgzf_strerror(r), b200bgzf_last_error(ctx)); ret = 1; }
            uint32_t bs = 1;
            b200bgzf_piece_spec sp;
            if (b200bgzf_container_plan(kind, param, &bs, &sp) == 0) units = (long)((n + bs - 1) / bs);
        }
    }
    if (!ret && produced && fwrite(dst, 1, produced, out) != produced) { fprintf(stderr, "%s: write error\n", name); ret = 1; }
    if (!ret) fprintf(stderr, "%ld done.\n", units);
    fflush(out);
    if (out != stdout) fclose(out);
    if (in != stdin) fclose(in);
    free(dst);
    free(src);
    if (multi) b200bgzf_multi_destroy(multi);
    else b200bgzf_destroy(ctx);
    return ret;
}
