"""Instruction-mnemonic histogram per kernel of the built library: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "7bgzf_b200", "lib7bgzf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
show = ["UBLKCP", "SYNCS", "LDS", "STS", "ATOMS", "LDG", "STG", "VOTE", "MATCH", "SHFL", "REDUX", "BAR", "WARPSYNC", "SHF", "LOP3", "IMAD", "IADD3",
        "ISETP", "BRA", "FLO", "BREV", "POPC", "NANOSLEEP", "HMMA", "UTCHMMA", "LDTM"]
print("# cuobjdump -sass 7bgzf_b200/lib7bgzf_b200.so : instruction mnemonic histogram per kernel (sm_100a)")
print("# Blackwell/Hopper-native evidence: UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier ops; VOTE/MATCH/SHFL/REDUX = warp collectives;")
print("# no HMMA/UTC*MMA/LDTM anywhere: nothing on this path is a dense contraction.")
name, hist = None, None
def flush():
    if name:
        n = sum(hist.values())
        print(f"\n## {name}  ({n} instructions)")
        print("  " + "  ".join(f"{k}={sum(v for m, v in hist.items() if m.split('.')[0] == k)}" for k in show
                               if k in ("UBLKCP", "HMMA", "UTCHMMA", "LDTM") or any(m.split('.')[0] == k for m in hist)))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, hist = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and hist is not None:
        hist[m.group(1)] += 1
flush()
