"""the other block-gzip containers (MiGz at any member size, GZinga, dictzip, RAZF) and
whole-stream gzip, written around raw DEFLATE pieces.  CPU part: the emulator makes the pieces (the very code the kernel
runs), the library's framing (7bgzf_b200/host/containers.c, no GPU involved) wraps them, and the reference's own decoders
(`7gzip -d`, `7migz -d`, `7gzinga -cd`, `7dictzip -cd`, `7razf -cd` of the compiled reference) must give the input back.
GPU part: the product's containers are byte-identical to the emulated ones."""
import ctypes
import gzip
import os
import struct
import zlib

import pytest

import helpers as H

B = H.codec_module()
needs_ref = pytest.mark.skipif(not os.path.exists(H.REF_CIELBOX), reason="oracle/_ref not built")
KINDS = {"gzip": B.CONTAINER_GZIP, "migz": B.CONTAINER_MIGZ, "gzinga": B.CONTAINER_GZINGA, "dictzip": B.CONTAINER_DICTZIP, "razf": B.CONTAINER_RAZF}
APPLET = {"gzip": "7gzip", "migz": "7migz", "gzinga": "7gzinga", "dictzip": "7dictzip", "razf": "7razf"}
INPUTS = {
    "fastq": H.synth("fastq", 700001), "sam": H.synth("sam", 300000), "noise": H.lcg_noise(200000), "one": b"x",
    "zeros": bytes(150000), "edge": H.synth("fastq", 2 * 51200),      # exactly one GZinga member, two pieces
}


def test_crc32_combine_matches_zlib():
    lib = B.load()
    a, b = H.lcg_noise(1000), H.synth("sam", 70001)
    for x, y in ((a, b), (b, a), (a, b""), (b"", a), (b, b)):
        assert lib.b200bgzf_crc32_combine(zlib.crc32(x), zlib.crc32(y), len(y)) == zlib.crc32(x + y)
    assert lib.b200bgzf_crc32_combine(zlib.crc32(a), zlib.crc32(bytes(1 << 20)), 1 << 20) == zlib.crc32(a + bytes(1 << 20))
    assert lib.b200bgzf_crc32_combine(0x12345678, 0x9abcdef0, 5 << 32) == H.oracle().oracle_crc32_combine(0x12345678, 0x9abcdef0, 5 << 32)


def test_plans_keep_a_stored_piece_inside_its_slot():
    lib = B.load()
    for kind in KINDS.values():
        for param in (0, 1, 63, 64, 100, 512, 1000, 4000):
            if kind == B.CONTAINER_DICTZIP and param and param < 16:
                continue
            safe = 0x80000000 if kind == B.CONTAINER_MIGZ else 0
            bs, sp = B.container_plan(kind, (param if kind in (B.CONTAINER_MIGZ, B.CONTAINER_DICTZIP) else 0) | safe, lib)
            assert 0 < bs <= 65280 and bs + 5 + sp.head_gap + sp.tail_gap <= 65536
            if kind == B.CONTAINER_MIGZ:
                assert bs * sp.member_blocks == (param or 512) * 1024
                bs, sp = B.container_plan(kind, param, lib)              # the first choice: pieces of up to 64 KiB
                assert 0 < bs <= 65536 and bs * sp.member_blocks == (param or 512) * 1024
    with pytest.raises(B.B200BgzfError):
        B.container_plan(B.CONTAINER_DICTZIP, 65281, lib)
    with pytest.raises(B.B200BgzfError):
        B.container_plan(9, 0, lib)


@pytest.mark.parametrize("level", [1, 6, 9, 12])
def test_piece_and_history_phases_do_not_depend_on_the_thread_order(level):
    """race check of the piece-mode and history code: the emulator runs the phases with the threads in forward, reverse and
    shuffled order (as tests/test_emulator.py does for members); the bytes must not change"""
    data = H.synth("fastq", 40000) + H.lcg_noise(500) + H.synth("sam", 40000)
    hist, payload = data[: 16320], data[16320 : 16320 + 49152]
    for history, final, head, tail in ((b"", True, 10, 8), (b"", False, 0, 0), (hist, False, 0, 0), (hist, True, 0, 8), (data[:272], False, 20, 0)):
        base = H.emul_piece(payload, level, head, tail, final, 0, history)
        for order in (1, 2):
            assert H.emul_piece(payload, level, head, tail, final, order, history) == base, (level, len(history), final, order)


def test_gap_bytes_of_a_slice_of_the_piece_stream():
    """b200bgzf_pieces_gap_bytes with piece_base / piece_total (a GPU shard's slice): head gaps of the members that begin in the
    slice, tail gaps of those that end in it — against a walk over the pieces"""
    import ctypes
    lib = B.load()
    for k in (1, 2, 3, 7, 0xFFFFFFFF):
        for total in (1, 2, 6, 7, 20):
            for base in range(0, total):
                for nb in range(0, total - base + 1):
                    spec = B.PieceSpec(k, 11, 8, 0, base, total, 0, 0)
                    want = sum((11 if gb % k == 0 else 0) + (8 if (gb + 1) % k == 0 or gb + 1 == total else 0) for gb in range(base, base + nb))
                    assert lib.b200bgzf_pieces_gap_bytes(nb * 1000, 1000, ctypes.byref(spec)) == want, (k, total, base, nb)


def test_non_final_pieces_end_on_a_byte_and_decode_alone():
    data = H.synth("fastq", 4 * 32768 + 17)
    whole = zlib.decompressobj(-15)
    got = b""
    blocks = [data[i : i + 32768] for i in range(0, len(data), 32768)]
    for i, b in enumerate(blocks):
        last = i + 1 == len(blocks)
        m, crc = H.emul_piece(b, 6, final=last)
        assert crc == zlib.crc32(b)
        one = zlib.decompressobj(-15)
        assert one.decompress(m) == b and one.eof == last
        if not last:
            assert m.endswith(b"\x00\x00\xff\xff")          # the empty stored block of a full flush
        got += whole.decompress(m)
    assert got == data and whole.eof


@pytest.mark.parametrize("level", [1, 6, 12])
@pytest.mark.parametrize("name", sorted(INPUTS))
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_emulated_containers_decode_with_zlib(kind, name, level):
    if level != 6 and name not in ("fastq", "noise"):
        pytest.skip("levels 1 and 12 on two corpora")
    data = INPUTS[name]
    blob = H.emul_container(KINDS[kind], data, level)
    if kind == "razf":
        assert zlib.decompressobj(31).decompress(blob) == data      # (the block index follows the member)
    else:
        assert gzip.decompress(blob) == data


@needs_ref
@pytest.mark.parametrize("name", sorted(INPUTS))
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_emulated_containers_decode_with_the_reference_applets(kind, name):
    data = INPUTS[name]
    blob = H.emul_container(KINDS[kind], data, 6)
    if kind == "gzinga" and len(blob) < 32768:
        pytest.skip("the reference's GZinga reader looks for the index in the last 32 KiB and fails on smaller files (applet/7gzinga.c:232-236)")
    rc, out = H.ref_applet_decode(APPLET[kind], blob)
    assert out == data, (rc, len(out))


@needs_ref
@pytest.mark.parametrize("kib", [16, 63, 64, 100, 512])
def test_migz_members_of_any_size_through_the_reference_reader(kib):
    data = H.synth("sam", 1200000)
    blob = H.emul_container(B.CONTAINER_MIGZ, data, 6, kib)
    # every member: subfield MZ carries the DEFLATE size, ISIZE the member's share of the input
    off, total = 0, 0
    while off < len(blob):
        assert blob[off : off + 16] == bytes.fromhex("1f8b08040000000000ff08004d5a0400")
        dsz = struct.unpack_from("<I", blob, off + 16)[0]
        isize = struct.unpack_from("<I", blob, off + 20 + dsz + 4)[0]
        assert isize == min(kib * 1024, len(data) - total)
        total += isize
        off += 20 + dsz + 8
    assert off == len(blob) and total == len(data)
    rc, out = H.ref_applet_decode("7migz", blob)
    assert out == data, rc


PRIMED, INDEPENDENT = 0x40000000, 0x20000000


@needs_ref
def test_primed_pieces_reach_into_the_input_before_them():
    """dictionary priming (what pigz does between its chunks; SURVEY 8f rank 4's "GPU route"): gzip by default, MiGz on request.
    The streams decode everywhere; they are smaller than those of independent pieces; a primed piece decodes with the
    history as its dictionary"""
    data = H.synth("sam", 700001)
    primed, indep = H.emul_container(B.CONTAINER_GZIP, data, 6), H.emul_container(B.CONTAINER_GZIP, data, 6, INDEPENDENT)
    assert gzip.decompress(primed) == data and gzip.decompress(indep) == data
    assert len(primed) < 0.985 * len(indep)
    rc, out = H.ref_applet_decode("7gzip", primed)
    assert out == data, rc
    ref = _ref_written("gzip", data, __import__("pathlib").Path(__import__("tempfile").mkdtemp()))
    assert len(primed) <= 1.03 * len(ref)                         # the reference's one libdeflate call over the whole file
    m = H.emul_container(B.CONTAINER_MIGZ, data, 6, 512 | PRIMED)
    assert gzip.decompress(m) == data and len(m) < len(H.emul_container(B.CONTAINER_MIGZ, data, 6, 512))
    rc, out = H.ref_applet_decode("7migz", m)
    assert out == data, rc
    # members stay independent: the second member decodes on its own
    dsz = struct.unpack_from("<I", m, 16)[0]
    second = m[20 + dsz + 8 :]
    assert gzip.decompress(second) == data[512 * 1024 :]
    # one piece with its history
    hist, payload = data[32768 - 32640 : 32768], data[32768 : 65536]
    piece, crc = H.emul_piece(payload, 6, final=False, history=hist)
    assert crc == zlib.crc32(payload) and zlib.decompressobj(-15, zdict=hist).decompress(piece) == payload


def test_container_layouts():
    data = H.synth("fastq", 300000)
    # GZinga: index member lists the end offset of every member
    z = H.emul_container(B.CONTAINER_GZINGA, data, 6)
    tail = z.rindex(bytes.fromhex("1f8b0810000000000 0ff".replace(" ", "")))
    comment = z[tail + 10 : z.index(b"\x00", tail + 10)].decode()
    ends = [int(e.split(":")[1]) for e in comment.split(";") if e]
    assert len(ends) == 3 and ends[-1] == tail and all(z[e : e + 4] == b"\x1f\x8b\x08\x10" for e in ends)
    assert z[tail:].endswith(b"\x00\x03\x00" + bytes(8))
    # dictzip: the RA table holds every chunk's compressed size; chunks decode on their own
    d = H.emul_container(B.CONTAINER_DICTZIP, data, 6)
    xlen, ver, chlen, chcnt = struct.unpack_from("<H", d, 10)[0], *struct.unpack_from("<HHH", d, 16)
    assert d[12:14] == b"RA" and ver == 1 and chlen == 58315 and chcnt == 6 and xlen == 10 + 2 * chcnt
    sizes = struct.unpack_from("<%dH" % chcnt, d, 22)
    pos = 22 + 2 * chcnt
    for i, sz in enumerate(sizes):
        assert zlib.decompressobj(-15).decompress(d[pos : pos + sz]) == data[i * chlen : (i + 1) * chlen]
        pos += sz
    assert d[pos : pos + 2] == b"\x03\x00" and struct.unpack_from("<II", d, pos + 2) == (zlib.crc32(data), len(data)) and pos + 10 == len(d)
    # RAZF: big-endian index after the trailer
    r = H.emul_container(B.CONTAINER_RAZF, data, 6)
    fsize, index_at = struct.unpack_from(">QQ", r, len(r) - 16)
    assert fsize == len(data) and r[:19] == bytes.fromhex("1f8b08040000000000030700") + b"RAZF\x01\x80\x00"
    nblk, bin0 = struct.unpack_from(">IQ", r, index_at)
    assert nblk == (len(data) + 32767) // 32768 - 1
    cells = struct.unpack_from(">%dI" % nblk, r, index_at + 12)
    for i, c in enumerate(cells):
        blk = zlib.decompressobj(-15).decompress(r[bin0 + c :], 32768)
        assert blk == data[(i + 1) * 32768 : (i + 2) * 32768]
    assert struct.unpack_from("<II", r, index_at - 8) == (zlib.crc32(data), len(data))


def test_dictzip_splits_into_members_of_32762_chunks():
    # tiny chunks make the member limit reachable: 40000 chunks of 64 bytes -> two members
    data = H.synth("sam", 40000 * 64)
    d = H.emul_container(B.CONTAINER_DICTZIP, data, 6, 64)
    assert gzip.decompress(d) == data
    counts, pos = [], 0
    while pos < len(d):
        chcnt = struct.unpack_from("<H", d, pos + 20)[0]
        counts.append(chcnt)
        pos += 22 + 2 * chcnt + sum(struct.unpack_from("<%dH" % chcnt, d, pos + 22)) + 10
    assert counts == [32762, 40000 - 32762] and pos == len(d)
    if os.path.exists(H.REF_CIELBOX):
        rc, out = H.ref_applet_decode("7dictzip", d)
        assert out == data, rc


def _ref_written(kind, data, tmp_path):
    """the container as the reference's own writer makes it (libdeflate level 6)"""
    import subprocess
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    if kind == "dictzip":
        dst = tmp_path / "out.dz"
        subprocess.run([H.REF_CIELBOX, "7dictzip", "-cl6", str(src), str(dst)], capture_output=True, check=True)
        return dst.read_bytes()
    if kind == "razf":
        return subprocess.run([H.REF_CIELBOX, "7razf", "-cl6", str(src)], capture_output=True, check=True).stdout
    with open(src, "rb") as f:
        return subprocess.run([H.REF_CIELBOX, APPLET[kind], "-cl6"], stdin=f, capture_output=True, check=True).stdout


def _check_units(kind, blob, data):
    """every unit of the reader's list decodes (zlib) to its slice of the data"""
    units, total = B.container_units(KINDS[kind], blob)
    assert total == len(data)
    pos = 0
    for in_off, in_len, hdr_len, out_len, piece in units:
        o = zlib.decompressobj(-15)
        got = o.decompress(blob[in_off + hdr_len : in_off + in_len], out_len) if piece else o.decompress(blob[in_off + hdr_len : in_off + in_len - 8])
        assert piece or o.eof
        assert len(got) == out_len and got == data[pos : pos + out_len], (kind, in_off)
        pos += out_len
    assert pos == len(data)
    return units


@pytest.mark.parametrize("kind", ["dictzip", "gzinga", "gzip", "razf"])
def test_readers_list_the_units_of_our_containers(kind):
    for name in ("fastq", "noise", "one", "edge"):
        _check_units(kind, H.emul_container(KINDS[kind], INPUTS[name], 6), INPUTS[name])


@needs_ref
@pytest.mark.parametrize("kind", ["dictzip", "gzinga", "gzip", "razf"])
def test_readers_list_the_units_of_reference_written_containers(kind, tmp_path):
    data = INPUTS["fastq"]
    units = _check_units(kind, _ref_written(kind, data, tmp_path), data)
    assert len(units) == {"dictzip": 13, "gzinga": 7, "gzip": 1, "razf": 22}[kind]


GOLDEN = {"dictzip": "ref.dz", "gzinga": "ref.gzinga", "gzip": "ref.gz", "razf": "ref.raz", "migz": "ref.migz"}


def _golden_input():
    return H.synth("fastq", 90000) + H.lcg_noise(3000) + H.synth("sam", 60000)      # (tests/golden/make_container_fixtures.py)


def _golden(kind):
    with open(os.path.join(H.GOLDEN, "containers", GOLDEN[kind]), "rb") as f:
        return f.read()


@pytest.mark.parametrize("kind", ["dictzip", "gzinga", "gzip", "razf"])
def test_readers_on_the_committed_reference_written_files(kind):
    units = _check_units(kind, _golden(kind), _golden_input())
    assert len(units) == {"dictzip": 3, "gzinga": 2, "gzip": 1, "razf": 5}[kind]


def test_committed_reference_files_are_what_the_reference_writes_today(tmp_path):
    if not os.path.exists(H.REF_CIELBOX):
        pytest.skip("oracle/_ref not built")
    for kind in ("dictzip", "razf", "gzinga", "gzip"):
        now, then = _ref_written(kind, _golden_input(), tmp_path), _golden(kind)
        if kind == "gzip":                          # (its header carries the time of the run)
            now, then = now[:4] + now[8:], then[:4] + then[8:]
        assert now == then, kind


@pytest.mark.parametrize("kind", ["dictzip", "gzinga", "razf"])
def test_readers_reject_damaged_indexes(kind):
    blob = H.emul_container(KINDS[kind], INPUTS["sam"], 6)
    for bad in (blob[:-1], blob[: len(blob) // 2], blob[1:], b"", b"\x1f\x8b" + bytes(40)):
        with pytest.raises(B.B200BgzfError):
            B.container_units(KINDS[kind], bad)
    if kind == "dictzip":
        dmg = bytearray(blob)
        dmg[22] ^= 0x40                                  # a chunk size that no longer adds up to the trailer
        with pytest.raises(B.B200BgzfError):
            B.container_units(KINDS[kind], bytes(dmg))
    if kind == "razf":
        dmg = bytearray(blob)
        dmg[-1] ^= 0x10                                  # index offset
        with pytest.raises(B.B200BgzfError):
            B.container_units(KINDS[kind], bytes(dmg))


def test_empty_input():
    assert gzip.decompress(H.emul_container(B.CONTAINER_GZIP, b"", 6)) == b""
    assert gzip.decompress(H.emul_container(B.CONTAINER_GZINGA, b"", 6)) == b""
    assert gzip.decompress(H.emul_container(B.CONTAINER_DICTZIP, b"", 6)) == b""
    assert H.emul_container(B.CONTAINER_MIGZ, b"", 6) == b""           # (the reference writes nothing either)


# ---------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def codec():
    c = B.Codec(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 6, 12])
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_gpu_containers_are_the_emulated_ones(codec, kind, level):
    for name in ("fastq", "noise", "one", "edge"):
        data = INPUTS[name]
        assert codec.container(KINDS[kind], data, level) == H.emul_container(KINDS[kind], data, level), (kind, name, level)


@pytest.mark.gpu
@pytest.mark.parametrize("level", [1, 6, 9, 12])
def test_gpu_primed_containers_are_the_emulated_ones(codec, level):
    data = H.synth("fastq", 400000) + H.lcg_noise(70000) + H.synth("sam", 300001)
    for kind, param in ((B.CONTAINER_GZIP, 0), (B.CONTAINER_GZIP, INDEPENDENT), (B.CONTAINER_MIGZ, 512 | PRIMED), (B.CONTAINER_MIGZ, 100 | PRIMED)):
        blob = codec.container(kind, data, level, param)
        assert blob == H.emul_container(kind, data, level, param), (kind, hex(param), level)
        assert gzip.decompress(blob) == data


@pytest.mark.gpu
def test_gpu_primed_pieces_across_batches_and_shards(codec):
    """the history of a batch's (shard's) first piece is the end of the previous batch (shard): same bytes however the stream is cut"""
    data = H.synth("sam", 48 << 20)                # 1536 pieces of 32 KiB: several host batches
    whole = codec.container(B.CONTAINER_GZIP, data, 6)
    o = zlib.decompressobj(31)
    assert zlib.crc32(o.decompress(whole)) == zlib.crc32(data) and o.eof
    m = B.MultiCodec([0, 0, 0])
    try:
        assert m.container(B.CONTAINER_GZIP, data, 6) == whole
    finally:
        m.close()
    small = data[: 3 << 20]
    assert codec.container(B.CONTAINER_GZIP, small, 6) == H.emul_container(B.CONTAINER_GZIP, small, 6)
    # the piece API with a history the block size does not divide evenly into the window
    spec = B.PieceSpec(0xFFFFFFFF, 0, 0, 0, 0, 0, 16320, 0)
    stream, off, crc = codec.compress_pieces(small, spec, 6, 20000)
    d = zlib.decompressobj(-15)
    assert d.decompress(stream) == small and d.eof
    with pytest.raises(B.B200BgzfError):
        codec.compress_pieces(small, B.PieceSpec(0xFFFFFFFF, 0, 0, 0, 0, 0, 1000, 0), 6, 20000)      # not a multiple of 272
    with pytest.raises(B.B200BgzfError):
        codec.compress_pieces(small, B.PieceSpec(0xFFFFFFFF, 0, 0, 0, 0, 0, 32640, 0), 6, 40000)     # piece + history > 64 KiB


@pytest.mark.gpu
def test_gpu_pieces_api(codec):
    data = H.synth("sam", 5 * 65280 + 123)
    spec = B.PieceSpec(3, 12, 8, 0)
    stream, off, crc = codec.compress_pieces(data, spec, 6, 65280)
    assert len(off) == 6 and off[0] == 0 and crc == [zlib.crc32(data[i * 65280 : (i + 1) * 65280]) for i in range(6)]
    # members of three pieces: 12 free bytes, one DEFLATE stream, 8 free bytes
    for first, end in ((0, off[3]), (3, len(stream))):
        m = stream[off[first] : end]
        assert m[:12] == bytes(12) and m[-8:] == bytes(8)
        o = zlib.decompressobj(-15)
        assert o.decompress(m[12:-8]) == data[first * 65280 : (first + 3) * 65280] and o.eof
    # 65536-byte pieces fit while they compress; one that does not is reported, not truncated
    stream, off, crc = codec.compress_pieces(data, B.PieceSpec(1, 20, 8, 0), 6, 65536)
    assert len(off) == 5
    with pytest.raises(B.B200BgzfError) as e:
        codec.compress_pieces(H.lcg_noise(3 * 65536), B.PieceSpec(1, 20, 8, 0), 6, 65536)
    assert e.value.code == B.E_NOFIT


@pytest.mark.gpu
def test_gpu_migz_falls_back_to_small_pieces_on_incompressible_data(codec):
    data = H.synth("fastq", 1 << 20) + H.lcg_noise(1 << 19) + H.synth("sam", 1 << 19)
    blob = codec.container(B.CONTAINER_MIGZ, data, 6)
    assert blob == H.emul_container(B.CONTAINER_MIGZ, data, 6, 0x80000000) and gzip.decompress(blob) == data
    assert codec.container(B.CONTAINER_MIGZ, data[: 1 << 20], 6) == H.emul_container(B.CONTAINER_MIGZ, data[: 1 << 20], 6)


@pytest.mark.gpu
def test_gpu_large_inputs_decode(codec):
    data = H.synth("fastq", 256 << 20)
    g = codec.container(B.CONTAINER_GZIP, data, 6)
    assert len(g) < 0.26 * len(data)
    o = zlib.decompressobj(31)
    assert zlib.crc32(o.decompress(g)) == zlib.crc32(data) and o.eof
    m = codec.container(B.CONTAINER_MIGZ, data[: 64 << 20], 6)
    assert gzip.decompress(m) == data[: 64 << 20]
    if os.path.exists(H.REF_CIELBOX):
        for kind in ("gzinga", "dictzip", "razf", "migz"):
            blob = codec.container(KINDS[kind], data[: 64 << 20], 6)
            rc, out = H.ref_applet_decode(APPLET[kind], blob)
            assert out == data[: 64 << 20], (kind, rc)


@pytest.mark.gpu
def test_gpu_dictzip_splits_into_members_of_32762_chunks(codec):
    # tiny chunks make the member limit reachable: 80000 chunks of 64 bytes -> three members
    data = H.synth("sam", 80000 * 64)
    d = codec.container(B.CONTAINER_DICTZIP, data, 6, 64)
    assert gzip.decompress(d) == data
    counts, pos = [], 0
    while pos < len(d):
        chcnt = struct.unpack_from("<H", d, pos + 20)[0]
        sizes = struct.unpack_from("<%dH" % chcnt, d, pos + 22)
        counts.append(chcnt)
        pos += 22 + 2 * chcnt + sum(sizes) + 10
    assert counts == [32762, 32762, 80000 - 2 * 32762] and pos == len(d)
    if os.path.exists(H.REF_CIELBOX):
        rc, out = H.ref_applet_decode("7dictzip", d)
        assert out == data, rc


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_gpu_reads_its_own_containers(codec, kind):
    for name in ("fastq", "sam", "noise", "one", "edge", "zeros"):
        data = INPUTS[name]
        assert codec.container_inflate(KINDS[kind], codec.container(KINDS[kind], data, 6)) == data, (kind, name)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kind", ["dictzip", "gzinga", "gzip", "migz", "razf"])
def test_gpu_reads_reference_written_containers(codec, kind, tmp_path):
    data = H.synth("sam", 3000017)
    assert codec.container_inflate(KINDS[kind], _ref_written(kind, data, tmp_path)) == data


@pytest.mark.gpu
def test_gpu_unit_list_errors(codec):
    data = INPUTS["fastq"]
    blob = codec.container(B.CONTAINER_DICTZIP, data, 6)
    units, total = B.container_units(B.CONTAINER_DICTZIP, blob)
    assert codec.inflate_units(blob, units, total) == data
    # a piece that claims more output than its blocks hold / a unit beyond the input
    u = list(units)
    u[2] = (u[2][0], u[2][1] - 9, u[2][2], u[2][3], 1)
    with pytest.raises(B.B200BgzfError) as e:
        codec.inflate_units(blob, u, total)
    assert e.value.code == B.E_FORMAT
    u = list(units)
    u[-1] = (len(blob), 100, 0, 10, 1)
    with pytest.raises(B.B200BgzfError) as e:
        codec.inflate_units(blob, u, total)
    assert e.value.code == B.E_FORMAT
    assert codec.inflate_units(blob, [], 0) == b""


@pytest.mark.gpu
def test_gpu_large_dictzip_and_razf_round_trip(codec):
    data = H.synth("fastq", 128 << 20)
    for kind in (B.CONTAINER_DICTZIP, B.CONTAINER_RAZF, B.CONTAINER_GZINGA):
        blob = codec.container(kind, data, 6)
        assert zlib.crc32(codec.container_inflate(kind, blob)) == zlib.crc32(data)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_gpu_applet_personalities(codec, kind, tmp_path):
    """the applet under the reference's other names, with the reference's command lines: same bytes as the library call;
    the reference's applet reads them; ours reads the reference's"""
    import subprocess
    data = H.synth("fastq", 1500000)
    exe = os.path.join(os.path.dirname(B.APPLET_PATH), APPLET[kind])
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    enc = tmp_path / "enc.bin"
    if kind == "dictzip":
        r = subprocess.run([exe, "-cl6", str(src), str(enc)], capture_output=True)
        blob = enc.read_bytes()
    elif kind == "razf":
        r = subprocess.run([exe, "-cl6", str(src)], capture_output=True)
        blob = r.stdout
    else:
        r = subprocess.run([exe, "-cl6"], input=data, capture_output=True)
        blob = r.stdout
    assert r.returncode == 0, r.stderr
    assert b"compression level = 6 (libdeflate)" in r.stderr and b"done." in r.stderr and b"ellapsed time" in r.stderr
    assert blob == codec.container(KINDS[kind], data, 6)
    enc.write_bytes(blob)
    # decode: file operand for the indexed containers (as the reference wants it), stdin for gzip / MiGz
    if kind in ("gzip", "migz"):
        d = subprocess.run([exe, "-d"], input=blob, capture_output=True)
    else:
        d = subprocess.run([exe, "-cd", str(enc)], capture_output=True)
    assert d.returncode == 0 and d.stdout == data, d.stderr
    if os.path.exists(H.REF_CIELBOX):
        rc, out = H.ref_applet_decode(APPLET[kind], blob)
        assert out == data, rc
        theirs = tmp_path / "ref.bin"
        theirs.write_bytes(_ref_written(kind, data, tmp_path))
        if kind in ("gzip", "migz"):
            d = subprocess.run([exe, "-d"], input=theirs.read_bytes(), capture_output=True)
        else:
            d = subprocess.run([exe, "-cd", str(theirs)], capture_output=True)
        assert d.returncode == 0 and d.stdout == data, d.stderr
    if kind == "razf":                                  # the pieces over three contexts: the same file
        x = subprocess.run([exe, "-cl6", "--devices=3", str(src)], env=dict(os.environ, CUDA_VISIBLE_DEVICES="0,0,0"), capture_output=True)
        assert x.returncode != 0 or x.stdout == blob
    if kind == "gzip":
        x = subprocess.run([exe, "-cl6", "--independent"], input=data, capture_output=True)
        assert x.returncode == 0 and x.stdout == codec.container(KINDS[kind], data, 6, INDEPENDENT)
    if kind == "migz":
        x = subprocess.run([exe, "-cl6", "--primed", "-b", "256"], input=data, capture_output=True)
        assert x.returncode == 0 and x.stdout == codec.container(KINDS[kind], data, 6, 256 | PRIMED)
    if kind == "dictzip":
        x = subprocess.run([exe, "-cl6", "-X", str(src), str(enc)], capture_output=True)
        assert x.returncode == 0 and enc.read_bytes() == codec.container(KINDS[kind], data, 6, 65280)
    # damaged input: an error, not a crash
    bad = tmp_path / "bad.bin"
    bad.write_bytes(blob[: len(blob) // 2])
    e = subprocess.run([exe, "-cd", str(bad)] if kind not in ("gzip", "migz") else [exe, "-d"], input=blob[: len(blob) // 2] if kind in ("gzip", "migz") else None, capture_output=True)
    assert e.returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(GOLDEN))
def test_gpu_reads_the_committed_reference_written_files(codec, kind):
    assert codec.container_inflate(KINDS[kind], _golden(kind)) == _golden_input()


@pytest.mark.gpu
def test_gpu_containers_over_several_contexts_are_the_single_gpu_ones(codec):
    """SURVEY 8e for the containers: piece ranges (whole members where there are several) over the contexts, joined and framed
    once: the same bytes for every number of shards (three contexts on one GPU here, as in the BGZF sharding test)"""
    data = H.synth("fastq", 3 * 1024 * 1024 + 12345) + H.synth("sam", 1 << 20)
    for ctxs in ([0, 0, 0], [0, 0, 0, 0, 0]):
        m = B.MultiCodec(ctxs)
        try:
            for kind in sorted(KINDS):
                assert m.container(KINDS[kind], data, 6) == codec.container(KINDS[kind], data, 6), (kind, len(ctxs))
            assert m.container(B.CONTAINER_DICTZIP, data[: 80000 * 64], 6, 64) == codec.container(B.CONTAINER_DICTZIP, data[: 80000 * 64], 6, 64)
            assert m.container(B.CONTAINER_GZIP, b"x", 6) == codec.container(B.CONTAINER_GZIP, b"x", 6)
            assert m.container(B.CONTAINER_GZIP, b"", 6) == codec.container(B.CONTAINER_GZIP, b"", 6)
            noisy = data[: 1 << 20] + H.lcg_noise(1 << 19)
            assert m.container(B.CONTAINER_MIGZ, noisy, 6) == codec.container(B.CONTAINER_MIGZ, noisy, 6)     # (the redo in small pieces)
        finally:
            m.close()


@pytest.mark.gpu
def test_gpu_7gzip_reads_streams_of_sized_members_in_parallel(codec):
    """`7gzip -d` on a BGZF or MiGz stream: the members carry their size, so the header walk finds them (one warp each)
    instead of one warp for the whole file"""
    import subprocess
    data = H.synth("fastq", 5 << 20)
    bgzf = codec.compress(data, 6)
    assert codec.container_inflate(B.CONTAINER_GZIP, bgzf) == data
    assert codec.container_inflate(B.CONTAINER_GZIP, codec.container(B.CONTAINER_MIGZ, data, 6)) == data
    exe = os.path.join(os.path.dirname(B.APPLET_PATH), "7gzip")
    d = subprocess.run([exe, "-d"], input=bgzf, capture_output=True)
    assert d.returncode == 0 and d.stdout == data and b"82 done." in d.stderr        # 81 members + the EOF marker


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_gpu_verify_flag_checks_the_crc_of_every_member(codec, kind):
    """B200BGZF_VERIFY on the containers: members of any size against their trailers (64 KiB tiles combined on the device),
    members made of pieces (dictzip, RAZF) from the CRCs of their pieces combined on the host.  A flipped byte inside a
    stored block changes the output without upsetting the decoder: only the check sees it"""
    data = H.synth("fastq", 300000) + H.lcg_noise(200000) + H.synth("sam", 300000)
    blob = codec.container(KINDS[kind], data, 6)
    assert codec.container_inflate(KINDS[kind], blob, B.VERIFY) == data
    at = blob.index(H.lcg_noise(200000)[70000:70032])            # raw bytes of a stored piece
    dmg = bytearray(blob)
    dmg[at + 5] ^= 0x20
    got = codec.container_inflate(KINDS[kind], bytes(dmg))        # no check: decodes, one byte off
    assert got != data and len(got) == len(data)
    with pytest.raises(B.B200BgzfError) as e:
        codec.container_inflate(KINDS[kind], bytes(dmg), B.VERIFY)
    assert e.value.code == B.E_CRC


@pytest.mark.gpu
def test_gpu_unit_crcs(codec):
    data = H.synth("sam", 400000)
    blob = codec.container(B.CONTAINER_RAZF, data, 6)
    units, total = B.container_units(B.CONTAINER_RAZF, blob)
    out, crcs = codec.inflate_units(blob, units, total, want_crc=True)
    assert out == data and crcs == [zlib.crc32(data[i : i + 32768]) for i in range(0, len(data), 32768)]
    z = codec.container(B.CONTAINER_GZINGA, data, 6)               # members of 100 KiB: two tiles each on the device
    units, total = B.container_units(B.CONTAINER_GZINGA, z)
    out, crcs = codec.inflate_units(z, units, total, want_crc=True)
    assert out == data and crcs == [zlib.crc32(data[i : i + 102400]) for i in range(0, len(data), 102400)]
