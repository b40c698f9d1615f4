"""First-contact GPU check: compress/inflate a few MiB through the C ABI, compare with zlib and the emulator."""
import ctypes, gzip, os, sys, time, zlib
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "7bgzf_b200"))
import b200bgzf

root = os.path.join(os.path.dirname(__file__), "..")
gen = ctypes.CDLL(os.path.join(root, "build", "libdatagen.so"))
gen.b200gen_fill.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
gen.b200gen_fill.restype = ctypes.c_size_t
emu = ctypes.CDLL(os.path.join(root, "build", "libemul.so"))

def synth(kind, n):
    buf = bytearray(n)
    gen.b200gen_fill(kind, 1 + kind, b200bgzf._addr(buf), n)
    return bytes(buf)

def emul_stream(data, level, bs=0xff00):
    out = bytearray()
    dst = ctypes.create_string_buffer(65536)
    dl = ctypes.c_uint32()
    for off in range(0, len(data), bs):
        blk = data[off:off + bs]
        rc = emu.bgemul_compress_block(blk, len(blk), level, 0, dst, ctypes.byref(dl))
        assert rc == 0, rc
        out += dst.raw[:dl.value]
    return bytes(out) + b200bgzf.EOF_BLOCK

c = b200bgzf.Codec(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8 << 20
for kind, name in ((0, "fastq"), (1, "sam")):
    data = synth(kind, n)
    for level in (1, 6, 9):
        t = time.time(); comp = c.compress(data, level); dt = time.time() - t
        ok_zlib = gzip.decompress(comp) == data
        em = emul_stream(data[: 4 * 0xff00], level)
        ok_emul = comp[: len(em) - 28] == em[:-28]
        t = time.time(); back = c.inflate(comp); dt2 = time.time() - t
        print(f"{name} L{level}: in={len(data)} out={len(comp)} ratio={len(comp)/len(data):.4f} zlib_decodes={ok_zlib} "
              f"first4blocks_eq_emulator={ok_emul} gpu_inflate_ok={back == data}  compress {len(data)/dt/1e9:.2f} GB/s inflate {len(data)/dt2/1e9:.2f} GB/s (host, cold)")
# edge cases
for name, d in (("empty", b""), ("A", b"A"), ("zeros", bytes(65280)), ("noise", os.urandom(65280)), ("ragged", os.urandom(1000) + bytes(70000))):
    comp = c.compress(d, 6)
    print(name, len(d), len(comp), gzip.decompress(comp) == d, c.inflate(comp) == d)
# inflate a zlib-made stream (static+dynamic blocks, multiple)
co = zlib.compressobj(6, zlib.DEFLATED, -15)
payload = synth(0, 60000)
raw = co.compress(payload) + co.flush()
member = bytes.fromhex("1f8b08040000000000ff0600424302 00".replace(" ", "")) + (len(raw) + 25).to_bytes(2, "little") + raw + zlib.crc32(payload).to_bytes(4, "little") + len(payload).to_bytes(4, "little")
print("zlib-made member inflates:", c.inflate(member + b200bgzf.EOF_BLOCK) == payload)
prof = c.profile(True, True)
big = synth(0, 64 << 20)
c.compress(big, 6)
prof = c.profile(False, True)
tot = sum(prof[:10]) or 1
names = ["load", "crc+census", "hash", "build-peers", "search", "accept+jump", "walk", "tally", "huffman", "sizes+emit"]
nblk = (64 << 20) / 0xff00
print("cycles/block %.0f  build-link %.0f" % (sum(prof) / nblk, prof[10] / nblk))
print("  ".join(f"{n}={p / nblk:.0f}" for n, p in zip(names, prof[:10])))
print("  detail: sort=%.0f trees=%.0f header=%.0f codes=%.0f | accept=%.0f jump=%.0f | sizes+scan+zero=%.0f emit=%.0f" % tuple(x / nblk for x in (prof[11], prof[12], prof[13], prof[8], prof[15], prof[5], prof[14], prof[9])))
print("  walk: clear+mark=%.0f a=%.0f b=%.0f c=%.0f" % tuple(x / nblk for x in (prof[16], prof[17], prof[18], prof[6])))
print("  sizes=%.0f scans=%.0f zero=%.0f" % tuple(x / nblk for x in (prof[19], prof[20], prof[14])))
print("launches", c.launches())
