"""BASELINE.json configs 3 and 4 at one GPU's share, for the record: python tools/config_runs.py [pieces]
  config 4: high-ratio class (level 12) over <pieces> x 1 GiB synthetic FASTQ (seeds 1..), device-resident + ratio,
            next to the reference's libdeflate12 on a 16 MiB sample of the same text (all host threads)
  config 3: level 6 over <pieces> x 1 GiB synthetic SAM (seeds 2..), end to end through the host-buffer C ABI"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import b200bgzf, helpers as H
pieces = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 1 << 30
c = b200bgzf.Codec(0)
s = torch.cuda.current_stream().cuda_stream
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(c.bound(n), dtype=torch.uint8, pin_memory=True)
d_out = torch.empty(c.bound(n), dtype=torch.uint8, device="cuda")
d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
# ---- config 4
tot_ms, tot_out, ok = 0.0, 0, True
for i in range(pieces):
    H._gen().b200gen_fill(0, 1 + i, host.data_ptr(), n)
    d_in = host.cuda()
    if i == 0:
        c.compress_device(d_in.data_ptr(), 64 << 20, d_out.data_ptr(), d_out.numel(), 12, stream=s)   # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    clen = c.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), 12, stream=s)
    e1.record(); torch.cuda.synchronize()
    tot_ms += e0.elapsed_time(e1); tot_out += clen
    c.inflate_device(d_out.data_ptr(), clen, d_back.data_ptr(), n, stream=s)
    ok = ok and bool(torch.equal(d_back, d_in))
line = {"config": "4: libdeflate12 class, %d GiB synthetic FASTQ, 1 B200, device-resident" % pieces, "GBps": round(pieces * n / tot_ms / 1e6, 3),
        "ratio": round(tot_out / (pieces * n), 4), "roundtrip_ok": ok}
if H.have_ref():
    sample = H.synth("fastq", 16 << 20)
    cores = os.cpu_count() or 1
    _, sizes, t = H.Ref(12).compress_stream(sample, threads=cores, keep=False)
    line["reference"] = {"sample_MiB": 16, "threads": cores, "MBps": round(len(sample) / t / 1e6, 2), "ratio": round(sum(sizes) / len(sample), 4)}
print(json.dumps(line))
# ---- config 3 (one GPU's share)
tot_s, tot_out = 0.0, 0
for i in range(pieces):
    H._gen().b200gen_fill(1, 2 + i, host.data_ptr(), n)
    if i == 0:
        c.compress_into(host.data_ptr(), n, h_out.data_ptr(), h_out.numel(), 6)                           # warm-up
    t0 = time.perf_counter()
    clen = c.compress_into(host.data_ptr(), n, h_out.data_ptr(), h_out.numel(), 6)
    tot_s += time.perf_counter() - t0; tot_out += clen
print(json.dumps({"config": "3 (one GPU's share): libdeflate6 class, %d GiB synthetic SAM, end to end (pinned host buffers)" % pieces,
                  "GBps": round(pieces * n / tot_s / 1e9, 3), "ratio": round(tot_out / (pieces * n), 4)}))
