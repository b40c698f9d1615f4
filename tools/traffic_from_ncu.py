"""profiles/<round>_traffic.json from the round's `ncu --set full` captures (bench.py picks the newest round's file):
  python tools/traffic_from_ncu.py r02 gpurun_out/prof_compress_r02.ncu-rep gpurun_out/ncu_compress_r02.log \
                                       gpurun_out/prof_inflate_r02.ncu-rep gpurun_out/ncu_inflate_r02.log
The logs are tools/prof_run.py's stdout under ncu ("ok <payload bytes> <stream bytes>" = the algorithmic bytes of the
captured launch); traffic = dram__bytes_read.sum + dram__bytes_write.sum of that one launch."""
import csv, json, os, re, subprocess, sys


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "dram__bytes_read.sum" in r)
    units, vals = rows[rows.index(hdr) + 1], rows[rows.index(hdr) + 2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        tot += float(vals[i].replace(",", "")) * mul
    return int(tot)


def alg_bytes(log):
    m = re.search(r"^ok (\d+) (\d+)", open(log).read(), re.M)
    return int(m.group(1)) + int(m.group(2))


tag, crep, clog, irep, ilog = sys.argv[1:6]
res = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture per kernel, divided by the algorithmic bytes "
               "(payload + stream, SURVEY 8d) of the captured launch; bench.py scales it to the launch it times"}
for name, rep, log in (("compress", crep, clog), ("inflate", irep, ilog)):
    if not os.path.exists(rep):
        continue
    d, a = dram_bytes(rep), alg_bytes(log)
    res[name] = {"capture": os.path.basename(rep), "dram_bytes": d, "algorithmic_bytes": a, "dram_per_algorithmic_byte": round(d / a, 4)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", f"{tag}_traffic.json")
json.dump(res, open(path, "w"), indent=1)
print(json.dumps(res))
