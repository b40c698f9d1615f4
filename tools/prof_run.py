"""Small driver for ncu captures: python tools/prof_run.py <compress|inflate> <MiB> <level> <kind> [iters]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import b200bgzf, helpers as H
mode, mib, level, kind = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 2
n = mib << 20
c = b200bgzf.Codec(0)
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
H._gen().b200gen_fill(0 if kind == "fastq" else 1, 1 if kind == "fastq" else 2, host.data_ptr(), n)
d_in = host.cuda()
d_out = torch.empty(c.bound(n), dtype=torch.uint8, device="cuda")
d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
clen = 0
for i in range(iters):
    clen = c.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), level, stream=s)
    if mode == "inflate":
        c.inflate_device(d_out.data_ptr(), clen, d_back.data_ptr(), n, stream=s)
torch.cuda.synchronize()
print("ok", n, clen)
