"""Small end-to-end exercise for compute-sanitizer: a few blocks through compress (levels 1, 6, 12) and inflate."""
import os, sys, zlib
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "tests")); sys.path.insert(0, os.path.join(root, "7bgzf_b200"))
import helpers as H, b200bgzf
c = b200bgzf.Codec(0)
data = H.synth("fastq", 3 * H.BLOCK + 777) + H.lcg_noise(70000) + bytes(3000) + b"A"
for level in (1, 6, 12):
    comp = c.compress(data, level)
    assert c.inflate(comp) == data
    members, st = c.compress_blocks([data[:65280], b"", b"x", data[65280:130000]], level)
    assert st == [0, 0, 0, 0]
ref = b"".join(H.zlib_member(data[o:o + H.BLOCK], 6) for o in range(0, len(data), H.BLOCK))
assert c.inflate(ref + H.EOF_BLOCK) == data
print("sanity ok")
