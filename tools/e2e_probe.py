"""End-to-end host-buffer timings (pinned host buffers, PCIe inside the timed region): python tools/e2e_probe.py [MiB]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import b200bgzf, helpers as H
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
c = b200bgzf.Codec(0)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
H._gen().b200gen_fill(0, 1, h_in.data_ptr(), n)
bound = c.bound(n)
h_out = torch.empty(bound, dtype=torch.uint8, pin_memory=True)
h_back = torch.empty(n, dtype=torch.uint8, pin_memory=True)
clen = c.compress_into(h_in.data_ptr(), n, h_out.data_ptr(), bound, 6)
for name, fn in (("compress", lambda: c.compress_into(h_in.data_ptr(), n, h_out.data_ptr(), bound, 6)),
                 ("inflate", lambda: c.inflate_into(h_out.data_ptr(), clen, h_back.data_ptr(), n))):
    fn()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print("%s e2e: best %.2f ms (%.1f GB/s), median %.2f ms (%.1f GB/s)" % (name, min(ts) * 1e3, n / min(ts) / 1e9, sorted(ts)[2] * 1e3, n / sorted(ts)[2] / 1e9))
print("roundtrip ok:", bool(torch.equal(h_back, h_in)))
