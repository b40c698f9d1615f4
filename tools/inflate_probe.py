"""Device-resident inflate timing: python tools/inflate_probe.py [MiB]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import b200bgzf, helpers as H
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
c = b200bgzf.Codec(0)
for kind, seed, name in ((0, 1, "fastq"), (1, 2, "sam")):
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    H._gen().b200gen_fill(kind, seed, host.data_ptr(), n)
    d_in = host.cuda()
    d_out = torch.empty(c.bound(n), dtype=torch.uint8, device="cuda")
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    clen = c.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), 6, stream=s)
    for _ in range(2): c.inflate_device(d_out.data_ptr(), clen, d_back.data_ptr(), n, stream=s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): c.inflate_device(d_out.data_ptr(), clen, d_back.data_ptr(), n, stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s: inflate %.2f ms  %.1f GB/s  ok=%s" % (name, ms, n / ms / 1e6, bool(torch.equal(d_back, d_in))))
