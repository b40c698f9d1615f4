"""CPU: pins the oracle restatement (oracle/bgzf_oracle.c) against the reference's golden vectors
(tests/golden, generated from the unmodified reference by make_golden.py) and, when oracle/_ref is present,
against the compiled reference itself."""
import json
import os
import struct
import zlib

import pytest

import helpers as H

KA = json.load(open(os.path.join(H.GOLDEN, "known_answers.json")))
MAL = json.load(open(os.path.join(H.GOLDEN, "malformed.json")))
needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (make -f oracle/Makefile.ref)")


def test_crc32_known_answers():
    o = H.oracle()
    for name, payload in (("A", b"A"), ("zeros65280", bytes(65280)), ("zeros65536", bytes(65536)),
                          ("noise65280", H.lcg_noise(65280)), ("acgt65280", H.acgt(65280))):
        assert "%08x" % o.oracle_crc32(0, payload, len(payload)) == KA[f"{name}_L6"]["crc"]
    assert o.oracle_crc32(0, b"", 0) == 0


def test_crc32_combine_is_concatenation():
    o = H.oracle()
    a, b = H.lcg_noise(1000), H.synth("fastq", 70001)
    ca, cb = zlib.crc32(a), zlib.crc32(b)
    assert o.oracle_crc32_combine(ca, cb, len(b)) == zlib.crc32(a + b)
    assert o.oracle_crc32_combine(ca, 0, 0) == ca


def test_eof_and_frame_bytes():
    import ctypes
    o = H.oracle()
    buf = ctypes.create_string_buffer(64)
    assert o.oracle_eof_block(buf) == 28 and buf.raw[:28].hex() == KA["eof_L6"]["hex"]
    # the reference stores "A" (passthrough): framing + stored block must reproduce its bytes exactly
    dst = ctypes.create_string_buffer(64)
    n = ctypes.c_size_t(64)
    assert o.oracle_store_deflate(dst, ctypes.byref(n), b"A", 1) == 0 and dst.raw[: n.value] == bytes.fromhex("010100feff41")
    member = ctypes.create_string_buffer(64)
    total = o.oracle_bgzf_frame(member, dst.raw[: n.value], n.value, b"A", 1)
    assert total == KA["A_L6"]["size"] and member.raw[:24].hex() == KA["A_L6"]["head"]
    assert o.oracle_passthrough(6) == 31 and o.oracle_passthrough(12) == 7


def test_store_deflate_two_blocks_and_capacity():
    import ctypes
    o = H.oracle()
    src = H.lcg_noise(65536)
    dst = ctypes.create_string_buffer(70000)
    n = ctypes.c_size_t(70000)
    assert o.oracle_store_deflate(dst, ctypes.byref(n), src, len(src)) == 0 and n.value == 65536 + 10
    assert zlib.decompress(dst.raw[: n.value], -15) == src
    n = ctypes.c_size_t(65540)
    assert o.oracle_store_deflate(dst, ctypes.byref(n), src, len(src)) != 0


def test_read_gz_header_variants():
    import ctypes
    o = H.oracle()
    eo, el, bl = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    m = H.zlib_member(b"hello hello hello")
    assert o.oracle_read_gz_header(m, len(m), eo, el, bl) == 18 and bl.value == len(m) and (eo.value, el.value) == (12, 6)
    migz = bytes.fromhex("1f8b08040000000000ff0800") + b"MZ\x04\x00" + struct.pack("<I", 100)
    assert o.oracle_read_gz_header(migz, len(migz), eo, el, bl) == 20 and bl.value == 100 + 20 + 8
    assert o.oracle_read_gz_header(b"\x1f\x8b\x08\x00" + bytes(20), 24, eo, el, bl) == 0   # plain gzip: no block length
    assert o.oracle_read_gz_header(b"PK\x03\x04" + bytes(20), 24, eo, el, bl) == 0


@pytest.mark.parametrize("kind", ["fastq", "sam"])
@pytest.mark.parametrize("level", [1, 6, 12])
def test_oracle_inflates_reference_golden_streams(kind, level):
    stream = open(os.path.join(H.GOLDEN, f"ref_{kind}_L{level}.bgz"), "rb").read()
    data = H.synth(kind, 2 * H.BLOCK)
    rc, out, nm = H.oracle_decompress(stream)
    assert rc == 0 and nm == 2 and out == data
    assert [m[1] for m in H.members(stream)] == KA[f"{kind}_L{level}_sizes"]
    for off, size, isize, crc in H.members(stream):
        pass
    assert H.gunzip(stream) == data


def test_oracle_inflate_rejects_the_reference_trees_malformed_vectors():
    """lib/isa-l/igzip/inflate_std_vects.h (SURVEY 8c negative fixture), framed as BGZF members by tests/golden/make_golden.py:
    the reference's decoder rejects all 151, so must the restatement"""
    cases = json.load(open(os.path.join(H.GOLDEN, "isal_std_vects.json")))["cases"]
    assert len(cases) == 151
    for c in cases:
        m = bytes.fromhex(c["hex"])
        isize = struct.unpack_from("<I", m, len(m) - 4)[0]
        rc, out = H.oracle_inflate_raw(m[18:-8], isize)
        assert (rc == 0 and len(out) == isize) == (c["ref_rc"] == 0), c["name"]


def test_oracle_inflate_matches_reference_verdict_on_malformed():
    """every corrupted member: same accept/reject as the reference's libdeflate decoder, same bytes when accepted"""
    n_ok = 0
    for c in MAL["cases"]:
        m = bytes.fromhex(c["hex"])
        isize = struct.unpack_from("<I", m, len(m) - 4)[0]
        rc, out = H.oracle_inflate_raw(m[18:-8], isize)
        ref_accepts = c["ref_rc"] == 0
        mine_accepts = rc == 0 and len(out) == isize
        assert mine_accepts == ref_accepts, (c["ref_rc"], rc, len(out), isize)
        if ref_accepts:
            assert "%08x" % zlib.crc32(out) == c["ref_out_crc"]
            n_ok += 1
    assert n_ok > 5


def test_oracle_inflates_zlib_streams_all_block_types():
    payload = H.synth("sam", 50000)
    for level, strat in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
        m = H.zlib_member(payload, level, strat)
        rc, out, nm = H.oracle_decompress(m + H.EOF_BLOCK)
        assert rc == 0 and out == payload and nm == 2


def test_parse_method_defaults():
    import ctypes
    o = H.oracle()
    lvl = ctypes.c_int()
    table = {None: (0, 6), b"": (0, 6), b"libdeflate": (5, 6), b"LibDeflate12": (5, 12), b"zlib9": (0, 9), b"slz": (4, 1), b"libslz1": (4, 1),
             b"7-zip": (1, 2), b"igzip3": (7, 3), b"zopfli": (2, 1), b"bogus7": (0, 7), b"zlibng": (6, 6), b"cryptopp": (8, 6), b"miniz": (3, 1)}
    for spec, want in table.items():
        assert (o.oracle_parse_method(spec, ctypes.byref(lvl)), lvl.value) == want, spec


@needs_ref
def test_reference_known_answers_reproduce():
    """the golden file is what the compiled reference produces here (guards against a stale fixture)"""
    for level in (1, 6, 12):
        for name, payload in (("A", b"A"), ("zeros65280", bytes(65280)), ("acgt65280", H.acgt(65280)), ("noise65280", H.lcg_noise(65280))):
            rc, member, _ = H.Ref(level).bgzf_compress(payload)
            assert rc == 0 and len(member) == KA[f"{name}_L{level}"]["size"] and member[:24].hex() == KA[f"{name}_L{level}"]["head"]
        assert H.Ref(level).bgzf_compress(b"", 28)[1].hex() == KA["eof_L6"]["hex"]
        assert H.Ref(level).bgzf_compress(b"", 27)[0] == -1


@needs_ref
def test_oracle_vs_reference_decoder_and_crc():
    data = H.synth("fastq", 300000) + H.lcg_noise(5000) + bytes(70000)
    ref = H.Ref(6)
    stream, sizes, _ = ref.compress_stream(data)
    rc, out, _ = ref.inflate_stream(stream)
    assert rc == 0 and out == data
    rc2, out2, nm = H.oracle_decompress(stream)
    assert rc2 == 0 and out2 == out and nm == len(sizes)
    for off, size, isize, crc in H.members(stream):
        pass
    assert ref.crc32(data[:65280]) == H.oracle().oracle_crc32(0, data[:65280], 65280) == H.members(stream)[0][3]
