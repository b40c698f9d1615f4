"""Per-function shares of an ncu report (executed warp-instructions and stall samples by the source function a line
belongs to): python tools/ncu_funcs.py <report.ncu-rep>"""
import csv, os, re, subprocess, sys
rep = sys.argv[1]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "7bgzf_b200", "csrc")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
funcs = {}
for fn in os.listdir(root):
    cur, table = None, []
    for i, line in enumerate(open(os.path.join(root, fn), errors="replace"), 1):
        m = re.match(r"^(?:BG_HD|__device__|__global__|static|extern|template)?.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$", line)
        if m and not line.startswith((" ", "\t", "#", "/", "*", "}")) and "(" in line:
            cur = m.group(1)
        table.append(cur)
    funcs[fn] = table
rows = list(csv.reader(out.splitlines()))
fname, hdr, agg = None, None, {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 7:
        d = dict(zip(hdr, r))
        try:
            ins, smp = int(d["Instructions Executed"]), int(d["# Samples"])
            thr = int(d.get("Thread Instructions Executed", "0") or 0)
        except ValueError:
            continue
        ln = int(r[0])
        f = funcs.get(fname, [])
        name = (f[ln - 1] if ln - 1 < len(f) else None) or fname
        a = agg.setdefault(f"{fname}:{name}", [0, 0, 0])
        a[0] += ins; a[1] += smp; a[2] += thr
ti, ts = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
print(f"{'function':58s} %instr %samples lanes/instr")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{k:58s} {100*a[0]/ti:6.2f} {100*a[1]/ts:7.2f} {a[2]/a[0] if a[0] else 0:6.1f}")
