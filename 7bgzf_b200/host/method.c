/*
 * method.c — BGZF_METHOD=<name><digits> parser (host C), mirroring bgzf_compress.c:53-113 of the reference:
 * trailing decimal digits are the level, the rest is a case-insensitive method name, unset/unknown means
 * zlib, and each method has its own default level.  Every method name is served by the one GPU codec
 * (there is no multi-backend dispatch), so only the level survives; levels outside 1..12 are rejected
 * instead of crashing (the reference dereferences NULL for libdeflate13).
 */
#include <ctype.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>

#include "../../include/b200bgzf.h"

struct method_def {
    const char *name;
    const char *canon;
    int default_level;
};

static const struct method_def k_methods[] = {
    { "zlib", "zlib", 6 },       { "7zip", "7zip", 2 },         { "7-zip", "7zip", 2 },
    { "zopfli", "zopfli", 1 },   { "miniz", "miniz", 1 },       { "slz", "slz", 1 },
    { "libslz", "slz", 1 },      { "libdeflate", "libdeflate", 6 }, { "zlibng", "zlibng", 6 },
    { "igzip", "igzip", 1 },     { "cryptopp", "cryptopp", 6 },
};

int b200bgzf_parse_method(const char *spec, int *level, char *method_name, size_t method_cap)
{
    const struct method_def *m = &k_methods[0]; /* unset or unknown name: zlib (bgzf_compress.c:54,103) */
    int lvl = -1;
    if (spec && *spec) {
        size_t l = strlen(spec), i = l;
        int digit = 1;
        while (i > 0 && isdigit((unsigned char)spec[i - 1])) {
            if (lvl < 0) lvl = 0;
            if (lvl < 100000) lvl += digit * (spec[i - 1] - '0');
            if (digit < 100000) digit *= 10;
            i--;
        }
        for (size_t k = 0; k < sizeof k_methods / sizeof k_methods[0]; k++)
            if (strlen(k_methods[k].name) == i && !strncasecmp(spec, k_methods[k].name, i)) {
                m = &k_methods[k];
                break;
            }
    }
    if (lvl < 0) lvl = m->default_level;
    if (method_name && method_cap) snprintf(method_name, method_cap, "%s", m->canon);
    if (lvl == 0) lvl = 1;              /* "level 0" of zlib-style methods: the cheapest class this codec has */
    if (lvl > 12) return -1;
    if (level) *level = lvl;
    return 0;
}
