"""Differential fuzz of the compress kernel against the CPU run of the same block algorithm (tests/model/emul.cpp):
random mixtures of text, runs, periodic data, noise and synthetic records, random block sizes and levels.
  python tools/compress_fuzz.py [cases] [seed]     (GPU box; prints one summary line)"""
import os, random, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "tests")); sys.path.insert(0, os.path.join(root, "7bgzf_b200"))
import helpers as H, b200bgzf
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
fq, sam, bam = H.synth("fastq", 1 << 20), H.synth("sam", 1 << 20), H.bamlike(1 << 19)
elf = open(os.path.join(root, "oracle", "_ref", "cielbox_ref"), "rb").read() if os.path.exists(os.path.join(root, "oracle", "_ref", "cielbox_ref")) else bam


def piece():
    k = rnd.randrange(9)
    n = rnd.choice([1, 3, 17, 100, 1000, 5000, 20000, 70000])
    if k == 0: return bytes(rnd.randrange(256) for _ in range(min(n, 4000)))
    if k == 1: return bytes([rnd.randrange(256)]) * n
    if k == 2:
        p = bytes(rnd.randrange(256) for _ in range(rnd.choice([2, 3, 4, 5, 7, 8, 31, 32, 33, 255, 258, 259])))
        return (p * (n // len(p) + 1))[:n]
    src = (fq, sam, bam, elf)[rnd.randrange(4)]
    o = rnd.randrange(max(1, len(src) - n))
    if k == 3:
        b = bytearray(src[o : o + n])
        for _ in range(max(1, n // 50)): b[rnd.randrange(len(b))] = rnd.randrange(256)
        return bytes(b)
    if k == 4: return bytes(rnd.choice(b"ACGT") for _ in range(min(n, 30000)))
    return src[o : o + n]


c = b200bgzf.Codec(0)
bad = 0
total = 0
for i in range(cases):
    data = b"".join(piece() for _ in range(rnd.randrange(1, 8)))[: 400000]
    level = rnd.choice([1, 2, 3, 4, 5, 6, 6, 6, 7, 8, 9, 10, 12]) if len(data) < 150000 else rnd.choice([1, 3, 6, 6, 7, 9])
    bs = rnd.choice([0xFF00, 0xFF00, 0x10000 - 64, 4099, 32768, 65535, 1000, 257])
    if len(data) // bs > 400: bs = 0xFF00
    got = c.compress(data, level, block_size=bs)
    want = H.emul_stream(data, level, block=bs)
    total += len(data)
    if got != want:
        bad += 1
        print(f"case {i}: level {level} block {bs} len {len(data)}: GPU stream differs from the emulator's")
    if H.gunzip(got) != data or c.inflate(got) != data:
        bad += 1
        print(f"case {i}: level {level} block {bs} len {len(data)}: round trip failed")
    if i % 3 == 0 and data:
        # piece mode with a random member layout: every slot as the emulator makes it, every member one DEFLATE stream
        import zlib
        k, hg, tg, nf = rnd.choice([1, 2, 3, 7, 0xFFFFFFFF]), rnd.randrange(0, 65), rnd.randrange(0, 65), rnd.randrange(4) == 0
        pbs = min(bs, 65536 - 5 - hg - tg)
        hcfg = rnd.choice([0, 0, 272, 4080, 16320, 32640])           # dictionary priming
        if hcfg: pbs = min(pbs, 65536 - hcfg)
        spec = b200bgzf.PieceSpec(k, hg, tg, 1 if nf else 0, 0, 0, hcfg, 0)
        stream, off, crc = c.compress_pieces(data, spec, level, pbs)
        blocks = [data[j : j + pbs] for j in range(0, len(data), pbs)]
        want = bytearray()
        for j, b in enumerate(blocks):
            first, last = j % k == 0, (j + 1) % k == 0 or j + 1 == len(blocks)
            h = min(hcfg, (j % k) * pbs)
            h -= h % 272
            m, cr = H.emul_piece(b, level, hg if first else 0, tg if last else 0, last and not nf, history=data[j * pbs - h : j * pbs])
            if off[j] != len(want) or crc[j] != cr:
                bad += 1
                print(f"case {i}: piece {j}: offset / CRC differs")
                break
            want += m
        if bytes(want) != stream:
            bad += 1
            print(f"case {i}: level {level} block {pbs} spec ({k},{hg},{tg},{nf}): piece stream differs from the emulator's")
        # (one member when k covers everything: the pieces between the gaps are one DEFLATE stream)
        if k == 0xFFFFFFFF:
            o = zlib.decompressobj(-15)
            if o.decompress(stream[hg : len(stream) - tg]) != data or o.eof == nf:
                bad += 1
                print(f"case {i}: the member's DEFLATE stream does not decode")
print(f"{cases} cases, {total >> 20} MiB: {bad} mismatches")
sys.exit(1 if bad else 0)
