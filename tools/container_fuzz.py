"""Mutation fuzz of the containers' read side on the GPU: python tools/container_fuzz.py [cases] [seed]
Random bit flips / byte changes / truncations in dictzip, RAZF, GZinga, MiGz and gzip files made by this codec and by the
reference; every read either fails cleanly (a b200bgzf error code) or returns bytes; with B200BGZF_VERIFY a read that
succeeds must return the original bytes (a CRC-32 collision aside).  Never a crash, a hang or a wrong "verified" result."""
import os, random, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "tests")); sys.path.insert(0, os.path.join(root, "7bgzf_b200"))
import helpers as H, b200bgzf as B
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 600
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
c = B.Codec(0)
data = H.synth("fastq", 150000) + H.lcg_noise(40000) + H.synth("sam", 120000)
kinds = {"gzip": B.CONTAINER_GZIP, "migz": B.CONTAINER_MIGZ, "gzinga": B.CONTAINER_GZINGA, "dictzip": B.CONTAINER_DICTZIP, "razf": B.CONTAINER_RAZF}
files = [(k, v, c.container(v, data, 6, 64 if k == "migz" else 0)) for k, v in kinds.items()]
gold = os.path.join(root, "tests", "golden", "containers")
gold_in = H.synth("fastq", 90000) + H.lcg_noise(3000) + H.synth("sam", 60000)
ref_files = [(k, kinds[k], open(os.path.join(gold, f), "rb").read()) for k, f in (("dictzip", "ref.dz"), ("razf", "ref.raz"), ("gzinga", "ref.gzinga"), ("migz", "ref.migz"))]
stats = {"clean_error": 0, "ok_same": 0, "ok_differs_unverified": 0, "crc_caught": 0, "bad": 0}
for i in range(cases):
    name, kind, blob = (files + ref_files)[i % (len(files) + len(ref_files))]
    want = data if i % (len(files) + len(ref_files)) < len(files) else gold_in
    b = bytearray(blob)
    for _ in range(rnd.choice([1, 1, 2, 5])):
        pos = rnd.randrange(len(b))
        b[pos] = rnd.randrange(256) if rnd.random() < 0.3 else b[pos] ^ (1 << rnd.randrange(8))
    if rnd.random() < 0.1:
        b = b[: rnd.randrange(1, len(b))]
    b = bytes(b)
    for flags in (0, B.VERIFY):
        try:
            got = c.container_inflate(kind, b, flags)
        except B.B200BgzfError as e:
            if e.code == B.E_CRC:
                stats["crc_caught"] += 1
            elif e.code in (B.E_FORMAT, B.E_NOSPACE, B.E_ARG):
                stats["clean_error"] += 1
            else:
                stats["bad"] += 1
                print(f"case {i} {name}: unexpected error {e}")
            continue
        if got == want:
            stats["ok_same"] += 1
        elif flags == 0:
            stats["ok_differs_unverified"] += 1
        else:
            # with the check on, a different result may only come from damage to bytes the check does not cover (a member's
            # ISIZE / index entry changing how much is decoded is caught by the size checks, not here)
            stats["bad"] += 1
            print(f"case {i} {name}: verified read returned different bytes ({len(got)} vs {len(want)})")
print(f"{cases} cases x 2 (plain, VERIFY): {stats}")
sys.exit(1 if stats["bad"] else 0)
