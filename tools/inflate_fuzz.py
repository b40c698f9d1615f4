"""Mutated members through the GPU inflate against the oracle decoder (same accept/reject, same bytes): python tools/inflate_fuzz.py [n]
One-off robustness run (the committed golden set has 120 such members with the reference decoder's verdicts)."""
import os, random, sys, zlib
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import b200bgzf, helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
c = b200bgzf.Codec(0)
rnd = random.Random(12345)
srcs = [H.synth("fastq", 50000), H.synth("sam", 65280), H.lcg_noise(3000) + bytes(4000) + H.synth("sam", 20000), b"x" * 70]
bases = []
for d in srcs:
    bases.append(c.compress(d, 6, eof=False))
    bases.append(c.compress(d, 1, eof=False))
    bases.append(H.zlib_member(d, 9))
    if H.have_ref():
        bases.append(H.Ref(6).bgzf_compress(d)[1])
agree = accept = reject = differ = 0
examples = []
for i in range(n):
    m = bytearray(rnd.choice(bases))
    for _ in range(rnd.choice((1, 1, 1, 2, 3))):
        k = rnd.randrange(18, len(m) - 8) if rnd.random() < 0.9 else rnd.randrange(len(m) - 8, len(m))   # deflate data, sometimes the trailer
        m[k] ^= 1 << rnd.randrange(8)
    m = bytes(m)
    rc, want, _ = H.oracle_decompress(m)
    isize = int.from_bytes(m[-4:], "little")
    try:
        got = c.inflate(m)
        g_ok = True
    except b200bgzf.B200BgzfError:
        got, g_ok = None, False
    # documented difference: output shorter than ISIZE is an error here, the reference emits the shorter output
    o_ok = rc == 0 and len(want) == isize
    if g_ok == o_ok and (not g_ok or got == want):
        agree += 1
        accept += g_ok
        reject += not g_ok
    else:
        differ += 1
        if len(examples) < 5: examples.append((i, rc, len(want) if want else None, isize, g_ok, None if got is None else len(got)))
print(f"{n} mutated members: {agree} agree ({accept} accepted with identical bytes, {reject} rejected), {differ} differ", examples)
