"""Compressed size of the block algorithm (CPU emulator = the GPU's bytes) against the compiled reference on text that is NOT the
benchmark corpus: a BAM-like binary record stream, C source, an ELF binary, word salad.  Needs oracle/_ref and /root/reference.
    python tools/experiments/ratio_probe.py"""
import os, sys, glob, struct, random
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, '7bgzf_b200'))
import helpers as H
def bamlike(n, seed=5):
    rnd = random.Random(seed); out = bytearray(); pos = 1000
    sam = H.synth("sam", n * 2)
    lines = [l for l in sam.split(b"\n") if l and not l.startswith(b"@")]
    code = {65:1, 67:2, 71:4, 84:8, 78:15}
    for l in lines:
        f = l.split(b"\t")
        if len(f) < 11: continue
        seq, qual, name = f[9], f[10], f[0] + b"\0"
        packed = bytes(((code.get(seq[i],15) << 4) | (code.get(seq[i+1],15) if i+1 < len(seq) else 0)) for i in range(0, len(seq), 2))
        q = bytes(max(0, c - 33) for c in qual)
        cigar = struct.pack("<I", (len(seq) << 4) | 0)
        core = struct.pack("<iiBBHHHIiii", 0, int(f[3]) & 0x7fffffff, len(name) & 255, int(f[4]) & 255, 4681, 1, int(f[1]) & 0xffff, len(seq), 0, int(f[7]) & 0x7fffffff, max(-2**31, min(2**31-1, int(f[8]))))
        tags = b"NMC" + bytes([rnd.randrange(4)]) + b"ASC" + bytes([150 - rnd.randrange(10)]) + b"RGZgrp1\0"
        rec = core + name + cigar + packed + q + tags
        out += struct.pack("<I", len(rec)) + rec
        if len(out) >= n: break
    return bytes(out[:n])
srcs = b"".join(open(f,'rb').read() for f in sorted(glob.glob('/root/reference/lib/libdeflate/*.c'))[:8])
bins = open('/usr/bin/python3.12','rb').read()[:4<<20] if os.path.exists('/usr/bin/python3.12') else b""
words = [w for w in open(os.path.join(ROOT, 'SURVEY.md'),'rb').read().split() if w]
rnd = random.Random(1); text = b" ".join(rnd.choice(words) for _ in range(600000))
cases = {"bam-like binary": bamlike(4<<20), "C source": srcs[:4<<20], "ELF binary": bins, "word salad": text[:4<<20]}
for name, data in cases.items():
    if len(data) < 100000: continue
    for level in (1, 6, 9):
        mine = len(H.emul_stream(data, level)) - 28
        ref = sum(H.Ref(level).compress_stream(data, keep=False)[1])
        print(f"{name:16s} L{level}: {len(data)>>10} KiB  ours {mine/len(data):.4f}  reference {ref/len(data):.4f}  delta {100*(mine/ref-1):+.2f}%")
