"""Regenerates tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile.ref from
/root/reference).  Run in the build container: python tests/golden/make_golden.py
Golden content:
  ref_<kind>_L<level>.bgz   the reference's bgzf_compress output for 2 blocks (2*0xff00 bytes) of synthetic data
  known_answers.json        sizes / CRC32 / return codes of the reference on the SURVEY 8(c) edge inputs
  malformed.json            corrupted DEFLATE payloads with the reference decoder's verdict (0 ok / non-zero error)
  isal_std_vects.json       the reference tree's own negative fixture (lib/isa-l/igzip/inflate_std_vects.h: 151 malformed raw
                            DEFLATE streams), each framed as a BGZF member, with the reference decoder's verdict
"""
import ctypes, json, os, random, sys, zlib
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import helpers as H

out = H.GOLDEN
ka = {}
for kind in ("fastq", "sam"):
    data = H.synth(kind, 2 * H.BLOCK)
    for level in (1, 6, 12):
        stream, sizes, _ = H.Ref(level).compress_stream(data)
        open(os.path.join(out, f"ref_{kind}_L{level}.bgz"), "wb").write(stream)
        ka[f"{kind}_L{level}_sizes"] = sizes

edge = {
    "A": b"A", "zeros65280": bytes(65280), "zeros65536": bytes(65536), "noise65280": H.lcg_noise(65280),
    "acgt65280": H.acgt(65280), "noise65536": H.lcg_noise(65536),
}
for level in (1, 6, 12):
    for name, payload in edge.items():
        rc, member, dl = H.Ref(level).bgzf_compress(payload)
        ka[f"{name}_L{level}"] = {"rc": rc, "size": len(member), "crc": "%08x" % zlib.crc32(payload),
                                  "head": member[:24].hex(), "payload_len": len(payload)}
    rc, m, dl = H.Ref(level).bgzf_compress(b"", 28); ka[f"eof_L{level}"] = {"rc": rc, "hex": m.hex()}
    ka[f"cap30_L{level}"] = {"rc": H.Ref(level).bgzf_compress(H.lcg_noise(1000), 30)[0]}
    ka[f"cap25_L{level}"] = {"rc": H.Ref(level).bgzf_compress(b"hello world", 25)[0]}
    ka[f"eofcap27_L{level}"] = {"rc": H.Ref(level).bgzf_compress(b"", 27)[0]}
json.dump(ka, open(os.path.join(out, "known_answers.json"), "w"), indent=1, sort_keys=True)

# malformed payloads: bit flips / truncations of valid streams, judged by the reference's libdeflate decoder
rng = random.Random(7)
lib = H.oracle()
ref = H.Ref(6)
dec_alloc = ctypes.CDLL(None)
cases = []
base_payload = H.synth("fastq", 6000)
for level, maker in ((6, lambda p: H.zlib_member(p, 6)), (1, lambda p: H.zlib_member(p, 1, zlib.Z_FIXED)), (9, lambda p: H.Ref(6).bgzf_compress(p)[1])):
    good = maker(base_payload)
    for i in range(40):
        m = bytearray(good)
        kind = rng.choice(["flip", "flip", "flip", "trunc"])
        if kind == "flip":
            pos = rng.randrange(18, len(m) - 8); m[pos] ^= 1 << rng.randrange(8)
        else:
            cut = rng.randrange(19, len(m) - 9)
            m = m[:cut] + m[-8:]
            m[16:18] = (len(m) - 1).to_bytes(2, "little")
        rc, outb, _ = ref.inflate_stream(bytes(m))
        cases.append({"hex": bytes(m).hex(), "ref_rc": rc, "ref_ok": rc == 0 and outb == base_payload, "ref_out_crc": "%08x" % zlib.crc32(outb)})
json.dump({"payload_len": len(base_payload), "cases": cases}, open(os.path.join(out, "malformed.json"), "w"))

# the reference's own malformed-stream vectors (parsed where they lie; only their bytes + the verdict are kept)
import re, struct
vects_h = "/root/reference/lib/isa-l/igzip/inflate_std_vects.h"
if os.path.exists(vects_h):
    text = open(vects_h).read()
    cases = []
    for name, body in re.findall(r"uint8_t\s+(std_vect_\d+)\[\]\s*=\s*\{([^}]*)\}", text):
        raw = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{1,2})", body))
        if len(raw) + 26 > 65536:
            continue
        # ISIZE: what the stream really inflates to when it is in fact decodable, else the largest payload
        try:
            d = zlib.decompressobj(-15)
            plain = d.decompress(raw, 65537)
            isize = len(plain) if d.eof and len(plain) <= 65536 else 65536
        except zlib.error:
            isize = 65536
        m = (bytes.fromhex("1f8b08040000000000ff060042430200") + struct.pack("<H", len(raw) + 25) + raw + struct.pack("<II", 0, isize))
        rc, outb, _ = ref.inflate_stream(m)
        cases.append({"name": name, "hex": m.hex(), "ref_rc": rc, "ref_out_crc": "%08x" % zlib.crc32(outb)})
    json.dump({"cases": cases}, open(os.path.join(out, "isal_std_vects.json"), "w"))
    print("isa-l vectors:", len(cases), "accepted by the reference:", sum(c["ref_rc"] == 0 for c in cases))
print("golden written:", sorted(os.listdir(out)))
