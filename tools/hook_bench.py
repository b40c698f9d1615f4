"""LD_PRELOAD hook throughput: N threads each calling bgzf_compress() with one 0xff00-byte block at a time
(the pattern of htslib's thread pool).  python tools/hook_bench.py <threads> <MiB> [level]"""
import ctypes, os, sys, threading, time
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "tests")); sys.path.insert(0, os.path.join(root, "7bgzf_b200"))
import helpers as H, b200bgzf
nth, mib = int(sys.argv[1]), int(sys.argv[2])
os.environ["BGZF_METHOD"] = "libdeflate" + (sys.argv[3] if len(sys.argv) > 3 else "6")
hook = ctypes.CDLL(b200bgzf.HOOK_PATH)
hook.bgzf_compress.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
data = H.synth("sam", mib << 20)
src = ctypes.create_string_buffer(data, len(data))
base = ctypes.addressof(src)
nb = (len(data) + H.BLOCK - 1) // H.BLOCK
out_sizes = [0] * nb
def work(tid):
    dst = ctypes.create_string_buffer(65536)
    for b in range(tid, nb, nth):
        n = ctypes.c_size_t(65536)
        ln = min(H.BLOCK, len(data) - b * H.BLOCK)
        rc = hook.bgzf_compress(dst, ctypes.byref(n), base + b * H.BLOCK, ln, 6)
        assert rc == 0
        out_sizes[b] = n.value
work(0) if nb < 0 else None
dst = ctypes.create_string_buffer(65536); n = ctypes.c_size_t(65536); hook.bgzf_compress(dst, ctypes.byref(n), base, H.BLOCK, 6)  # warm-up / init
t0 = time.time()
th = [threading.Thread(target=work, args=(t,)) for t in range(nth)]
[t.start() for t in th]; [t.join() for t in th]
dt = time.time() - t0
print(f"hook: {nth} threads, {mib} MiB SAM-like, {len(data)/dt/1e6:.1f} MB/s, ratio {sum(out_sizes)/len(data):.4f}, {dt/nb*nth*1e6:.0f} us per call per thread")
