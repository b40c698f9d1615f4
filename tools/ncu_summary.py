"""Write the judged summary of an ncu report: python tools/ncu_summary.py <report.ncu-rep> <out.txt> [launch csv]"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
keys = ["Duration", "Executed Ipc Active", "Executed Ipc Elapsed", "Issue Slots Busy", "SM Busy", "Memory Throughput", "DRAM Throughput", "L1/TEX Hit Rate",
        "L2 Hit Rate", "Avg. Active Threads Per Warp", "Executed Instructions ", "Registers Per Thread", "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block",
        "Theoretical Occupancy", "Achieved Occupancy", "Block Limit", "Grid Size", "Block Size", "Shared Memory Configuration Size", "No Eligible", "Warp Cycles Per Issued",
        "Mem Busy", "Max Bandwidth", "Mem Pipes Busy"]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on  ({rep.split('/')[-1]})\n")
    for line in det.splitlines():
        if any(k in line for k in keys) or "bgzf_" in line:
            f.write(line.rstrip() + "\n")
    f.write("\n# raw metrics\n")
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) >= 3:
        hdr, units, vals = rows[0], rows[1], rows[2]
        for h, u, v in zip(hdr, units, vals):
            if any(k in h for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput", "sm__warps_active.avg.pct",
                                    "launch__registers_per_thread", "sm__pipe_tensor", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_lsu",
                                    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct", "lts__t_bytes.sum ", "sm__cycles_elapsed.max")):
                f.write(f"{h} = {v} {u}\n")
    f.write("\n# hottest source lines (share of executed warp-instructions / of stall samples)\n")
    f.write(subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "25"], capture_output=True, text=True).stdout)
print(open(out).read()[:3000])
