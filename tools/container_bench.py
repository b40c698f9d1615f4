"""The other containers next to the reference's applets: python tools/container_bench.py [MiB] [ref MiB]
End to end from / to pinned host buffers through b200bgzf_container_compress_host / _inflate_host (H2D, kernels, D2H and the host
framing inside the timed region), median of 3 after a warm-up; the reference: its own applet (`7gzip`, `7migz`, `7gzinga`,
`7dictzip`, `7razf` -l6, -@ all cores) over files on /dev/shm, wall clock of the process.  One JSON line per container."""
import ctypes, json, os, subprocess, sys, time, zlib
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import b200bgzf as B, helpers as H

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ref_mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = mib << 20
c = B.Codec(0)
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
H._gen().b200gen_fill(0, 1, host.data_ptr(), n)
crc = zlib.crc32(host.numpy().tobytes())
cores = os.cpu_count()
PRIMED, INDEPENDENT = 0x40000000, 0x20000000
KINDS = (("gzip", B.CONTAINER_GZIP, "7gzip", 0), ("gzip-independent", B.CONTAINER_GZIP, "7gzip", INDEPENDENT), ("migz", B.CONTAINER_MIGZ, "7migz", 0),
         ("migz-primed", B.CONTAINER_MIGZ, "7migz", PRIMED), ("gzinga", B.CONTAINER_GZINGA, "7gzinga", 0),
         ("dictzip", B.CONTAINER_DICTZIP, "7dictzip", 0), ("razf", B.CONTAINER_RAZF, "7razf", 0))


def timed(f, reps=3):
    f()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2]


ref_cache = {}
for name, kind, applet, param in KINDS:
    cap = c.lib.b200bgzf_container_bound(kind, param, n)
    out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    back = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    got = ctypes.c_size_t()
    def comp():
        rc = c.lib.b200bgzf_container_compress_host(c.h, kind, param, host.data_ptr(), n, 6, out.data_ptr(), cap, ctypes.byref(got))
        assert rc == 0, rc
    tc = timed(comp)
    clen = got.value
    def dec():
        rc = c.lib.b200bgzf_container_inflate_host(c.h, kind, out.data_ptr(), clen, back.data_ptr(), n, ctypes.byref(got), 0)
        assert rc == 0 and got.value == n, (rc, got.value)
    line = {"container": name, "MiB": mib, "level": 6, "compress_e2e_GBps": round(n / tc / 1e9, 2), "ratio": round(clen / n, 4)}
    if kind != B.CONTAINER_GZIP:              # (a gzip member without an index is one unit: one warp — timed on a small sample below)
        td = timed(dec)
        assert zlib.crc32(back.numpy().tobytes()) == crc
        line["inflate_e2e_GBps"] = round(n / td / 1e9, 2)
    else:
        small = 8 << 20
        blob = c.container(kind, host.numpy()[:small].tobytes(), 6, param)
        t = time.perf_counter(); o = c.container_inflate(kind, blob); td = time.perf_counter() - t
        assert o == host.numpy()[:small].tobytes()
        line["inflate_one_warp_MBps"] = round(small / td / 1e6, 1)
    if applet in ref_cache:
        line["reference"] = ref_cache[applet]
    elif os.path.exists(H.REF_CIELBOX):
        rn = ref_mib << 20
        src, dst = "/dev/shm/cb_in.bin", "/dev/shm/cb_out.bin"
        with open(src, "wb") as f:
            f.write(host.numpy()[:rn].tobytes())
        th = [] if applet == "7gzip" else ["-@", str(cores)]          # (7gzip is one libdeflate call: no thread option)
        t = time.perf_counter()
        if name == "dictzip":
            subprocess.run([H.REF_CIELBOX, applet, "-cl6", *th, src, dst], capture_output=True, check=False)
        elif name == "razf":
            with open(dst, "wb") as fo:
                subprocess.run([H.REF_CIELBOX, applet, "-cl6", *th, src], stdout=fo, stderr=subprocess.DEVNULL, check=False)
        else:
            with open(src, "rb") as fi, open(dst, "wb") as fo:
                subprocess.run([H.REF_CIELBOX, applet, "-cl6", *th], stdin=fi, stdout=fo, stderr=subprocess.DEVNULL, check=False)
        rt = time.perf_counter() - t
        rsize = os.path.getsize(dst)
        t = time.perf_counter()
        with open(dst, "rb") as fi, open("/dev/null", "wb") as fo:
            if applet in ("7gzip", "7migz"):
                subprocess.run([H.REF_CIELBOX, applet, "-d", *th], stdin=fi, stdout=fo, stderr=subprocess.DEVNULL, check=False)
            else:
                subprocess.run([H.REF_CIELBOX, applet, "-cd", *th, dst], stdout=fo, stderr=subprocess.DEVNULL, check=False)
        rd = time.perf_counter() - t
        if name == "gzinga":
            # the reference's reader looks for the index in the last 32 KiB (applet/7gzinga.c:232-236): files of more than
            # about 2000 members cannot be read back by it; time its reader on a file that small
            rn2 = 128 << 20
            with open(src, "wb") as f:
                f.write(host.numpy()[:rn2].tobytes())
            with open(src, "rb") as fi, open(dst, "wb") as fo:
                subprocess.run([H.REF_CIELBOX, applet, "-cl6", *th], stdin=fi, stdout=fo, stderr=subprocess.DEVNULL)
            t = time.perf_counter()
            with open("/dev/null", "wb") as fo:
                ok = subprocess.run([H.REF_CIELBOX, applet, "-cd", *th, dst], stdout=fo, stderr=subprocess.DEVNULL).returncode == 0
            rd = (time.perf_counter() - t) * (rn / rn2) if ok else float("nan")
        line.update({"reference": {"MiB": ref_mib, "cores": 1 if applet == "7gzip" else cores, "compress_GBps": round(rn / rt / 1e9, 3), "inflate_GBps": round(rn / rd / 1e9, 3),
                                   "ratio": round(rsize / rn, 4)}})
        ref_cache[applet] = line["reference"]
        os.unlink(src); os.unlink(dst)
    print(json.dumps(line), flush=True)
