"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py <report.ncu-rep> [top]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 7:
        d = dict(zip(hdr, r))
        try:
            lines.append((int(d["Instructions Executed"]), int(d["# Samples"]), fname, int(r[0]), r[1].strip()[:100], d))
        except ValueError:
            pass
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[1] for l in lines) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("  %instr  %samples  file:line  source")
for i, s, f, ln, src, d in sorted(lines, key=lambda x: -x[1])[:top]:
    stalls = sorted(((int(v), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0), reverse=True)[:3]
    st = " ".join(f"{k[6:]}={v}" for v, k in stalls)
    print(f"  {100*i/tot_i:6.2f}  {100*s/tot_s:6.2f}  {f}:{ln}  {src}   [{st}]")
