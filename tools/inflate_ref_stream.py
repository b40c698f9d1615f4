"""Device-resident inflate of a stream made by the REFERENCE's libdeflate 6 (two dynamic blocks per member, 3-byte matches):
python tools/inflate_ref_stream.py [MiB]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import b200bgzf, helpers as H
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 512) << 20
c = b200bgzf.Codec(0)
for kind in ("fastq", "sam"):
    data = H.synth(kind, n)
    stream, sizes, t = H.Ref(6).compress_stream(data, threads=os.cpu_count() or 1)
    stream += H.EOF_BLOCK
    d_in = torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda()
    d_out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2): m = c.inflate_device(d_in.data_ptr(), len(stream), d_out.data_ptr(), d_out.numel(), stream=s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): c.inflate_device(d_in.data_ptr(), len(stream), d_out.data_ptr(), d_out.numel(), stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ok = bytes(d_out[:m].cpu().numpy()) == data
    print(f"{kind}: reference stream ratio {len(stream)/n:.4f} (made at {n/t/1e9:.2f} GB/s on {os.cpu_count()} threads); GPU inflate {ms:.2f} ms = {n/ms/1e6:.1f} GB/s, bit-exact {ok}")
