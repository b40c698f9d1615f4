"""Device-resident throughput and ratio per level: python tools/level_sweep.py [MiB]  (prints one JSON line per level and corpus)"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import b200bgzf, helpers as H
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
c = b200bgzf.Codec(0)
s = torch.cuda.current_stream().cuda_stream
for kind, seed, name in ((0, 1, "fastq"), (1, 2, "sam")):
    for level in (1, 3, 6, 9, 10, 12):
        n = (mib if level < 10 else max(mib // 4, 16)) << 20
        host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        H._gen().b200gen_fill(kind, seed, host.data_ptr(), n)
        d_in = host.cuda()
        d_out = torch.empty(c.bound(n), dtype=torch.uint8, device="cuda")
        d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
        clen = c.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), level, stream=s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if level < 10 else 1
        e0.record()
        for _ in range(reps): clen = c.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), level, stream=s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        c.inflate_device(d_out.data_ptr(), clen, d_back.data_ptr(), n, stream=s)
        print(json.dumps({"corpus": name, "level": level, "MiB": n >> 20, "compress_GBps": round(n / ms / 1e6, 2), "ratio": round(clen / n, 4),
                          "roundtrip_ok": bool(torch.equal(d_back, d_in))}))
