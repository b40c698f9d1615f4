#!/usr/bin/env python
"""bench.py — BGZF compress & inflate GB/s (uncompressed) on B200, next to the reference's CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mib M] [--level L] [--kind fastq|sam]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (N > 1)
  python bench.py --impl reference ...      the reference's own CPU implementation (oracle/_ref), all host threads

One "step" = one pass of the hot path over one batch: the whole workload (default 1 GiB of synthetic
FASTQ-like text cut into 0xff00-byte BGZF blocks, BGZF_METHOD=libdeflate6 class) compressed into one BGZF
stream; the inflate leg inflates that stream.  Every rank works on its own full-size batch (weak scaling);
blocks are independent, so there is no data-path collective — only the barrier and the max-over-ranks
reduction of the timing.
  value      device-resident compress throughput: input and output stay in HBM, CUDA events on the launch stream
  e2e        the same through the host-buffer C-ABI call (pinned host memory; H2D and D2H inside the timed region)
  inflate    the same two numbers for BGZF inflate of the stream just produced
  roofline   algorithmic bytes (payload read + stream written) / device time, against the measured HBM peak
  cpu_baseline  the reference (oracle/_ref: the unmodified 7bgzf hook + libdeflate) on this box's host cores
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "7bgzf_b200"))
BLOCK = 0xFF00


def load_gen():
    p = os.path.join(ROOT, "build", "libdatagen.so")
    if not os.path.exists(p):
        subprocess.run(["make", "-s", "testlibs"], cwd=ROOT, check=True)
    lib = ctypes.CDLL(p)
    lib.b200gen_fill.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t]
    lib.b200gen_fill.restype = ctypes.c_size_t
    return lib


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.15)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def cpu_reference(args, data_addr, nbytes, kind_name, steps=1, warmup=0):
    """the reference's CPU path (unmodified hook + libdeflate, oracle/_ref) on a bounded sample, all host threads"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    if not H.have_ref():
        return None
    cores = os.cpu_count() or 1
    # ~40 MB/s/core at level 6: size the sample for roughly 10-20 s of CPU work in total, at most the workload
    per_core = {1: 150e6, 6: 38e6, 9: 9e6, 12: 1.2e6}.get(args.level, 38e6 if args.level < 8 else 4e6)
    sample = int(min(nbytes, max(8 * BLOCK * cores, per_core * cores * 1.5)))
    sample -= sample % BLOCK
    ref = H.Ref(args.level)
    nb = sample // BLOCK
    sizes = (ctypes.c_uint32 * nb)()
    rc = ctypes.c_int()
    times = []
    # untimed warm-up on a small slice: the reference mallocs/frees its ~0.7 MB compressor per block, and the first
    # multi-threaded pass pays for the allocator's arena set-up
    H.oracle().refh_compress(ref.h, data_addr, min(sample, 16 * BLOCK * cores), BLOCK, cores, None, ctypes.addressof(sizes), ctypes.byref(rc))
    for i in range(warmup + steps):
        t = H.oracle().refh_compress(ref.h, data_addr, sample, BLOCK, cores, None, ctypes.addressof(sizes), ctypes.byref(rc))
        if i >= warmup:
            times.append(t)
    t = statistics.mean(times)
    out_bytes = sum(sizes)
    res = {"value": sample / t / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
           "sample": f"first {sample >> 20} MiB of the workload, reference bgzf_compress (BGZF_METHOD=libdeflate{args.level}) from {cores} pthreads over contiguous block ranges, memory to memory",
           "ratio": out_bytes / sample, "seconds": t}
    # inflate leg of the baseline: the reference's libdeflate decoder over the members it just could have written
    stream, _, _ = ref.compress_stream(ctypes.string_at(data_addr, min(sample, 64 << 20)), threads=cores)
    src = ctypes.create_string_buffer(stream, len(stream))
    out = ctypes.create_string_buffer(min(sample, 64 << 20))
    n, rc2 = ctypes.c_size_t(), ctypes.c_int()
    ti = min(H.oracle().refh_inflate(ref.h, ctypes.addressof(src), len(stream), cores, ctypes.addressof(out), len(out), ctypes.byref(n), ctypes.byref(rc2))
             for _ in range(3))
    res["inflate_value"] = n.value / ti / 1e9
    res["inflate_sample"] = f"{n.value >> 20} MiB, reference libdeflate_deflate_decompress per member from {cores} pthreads"
    return res


def bind_to_gpu_numa_node(index):
    """Multi-GPU boxes have two sockets: run this rank (and first-touch its pinned buffers) on the cores next to its
    GPU, or every host<->device copy of the end-to-end legs crosses the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1 and 64 * i + b < ncpu}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mib", type=int, default=1024, help="payload MiB per rank (BASELINE: 1 GiB)")
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--kind", default="fastq", choices=["fastq", "sam"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nbytes = args.mib << 20
    config = {"workload": f"{args.mib} MiB synthetic {args.kind.upper()}-like text (SURVEY App. B, seed {1 if args.kind == 'fastq' else 2}) per GPU, "
                          f"BGZF_METHOD=libdeflate{args.level} class, 0xff00-byte blocks; compress is the headline value, inflate of the same stream reported beside it",
              "block_bytes": BLOCK, "level": args.level, "bytes_per_gpu": nbytes, "parallelism": f"block-range sharding x{world}, no collective",
              "l2": f"inputs ({args.mib} MiB) are larger than the 126 MB L2; no explicit flush"}

    gen = load_gen()
    if args.impl == "reference":
        # the reference's own CPU implementation; rank 0 only
        if rank != 0:
            return
        sample_bytes = min(nbytes, 256 << 20)
        buf = ctypes.create_string_buffer(sample_bytes)
        gen.b200gen_fill(0 if args.kind == "fastq" else 1, 1 if args.kind == "fastq" else 2, buf, sample_bytes)
        r = cpu_reference(args, ctypes.addressof(buf), sample_bytes, args.kind, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/7bgzf_ref.so missing (run: make -f oracle/Makefile.ref)"}))
            return
        line = {"metric": "BGZF compress GB/s (uncompressed)", "value": r["value"], "unit": "GB/s", "n_gpus": args.gpus, "steps": max(1, args.steps),
                "warmup": min(args.warmup, 1), "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "inflate": {"value": r["inflate_value"], "unit": "GB/s", "sample": r["inflate_sample"]}, "ratio": r["ratio"]}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import b200bgzf

    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if numa:
        config["host_binding"] = f"each rank pinned to the {numa} cores of its GPU's NUMA node"
    if world > 1:
        # NCCL prints its version (and, with NCCL_DEBUG=INFO, much more) on stdout when the communicator comes up; stdout is
        # for the ONE JSON line, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    codec = b200bgzf.Codec(local_rank)
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    gen.b200gen_fill(0 if args.kind == "fastq" else 1, 1 if args.kind == "fastq" else 2, h_in.data_ptr(), nbytes)
    bound = codec.bound(nbytes)
    h_out = torch.empty(bound, dtype=torch.uint8, pin_memory=True)
    h_back = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = h_in.cuda()
    d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    d_back = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def run_timed(fn, steps, warmup, sampler_index=None):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = codec.launches()
        t0 = time.perf_counter()
        for a, b in evs:
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        launches = codec.launches() - l0
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        return dev_ms / 1e3, wall, launches

    res = {}
    with ClockSampler(local_rank) as clk:
        # --- compress, device resident: input/output stay in HBM; events on the stream the kernels run on
        clen = [0]

        def comp_dev():
            clen[0] = codec.compress_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), bound, args.level, stream=stream.cuda_stream)

        t_dev, _, launches_c = run_timed(comp_dev, args.steps, args.warmup)
        # --- inflate, device resident (member index built on the device inside the timed call)
        def inf_dev():
            codec.inflate_device(d_out.data_ptr(), clen[0], d_back.data_ptr(), nbytes, stream=stream.cuda_stream)

        t_idev, _, launches_i = run_timed(inf_dev, args.steps, args.warmup)
        # --- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
        def comp_e2e():
            clen[0] = codec.compress_into(h_in.data_ptr(), nbytes, h_out.data_ptr(), bound, args.level)

        _, w_e2e, _ = run_timed(comp_e2e, args.steps, args.warmup)

        def inf_e2e():
            codec.inflate_into(h_out.data_ptr(), clen[0], h_back.data_ptr(), nbytes)

        _, w_ie2e, _ = run_timed(inf_e2e, args.steps, args.warmup)
    ok = bool(torch.equal(d_back, d_in)) and bool(torch.equal(h_back, h_in))

    steps = args.steps
    t_dev_m, t_idev_m = max_over_ranks(t_dev), max_over_ranks(t_idev)
    w_e2e_m, w_ie2e_m = max_over_ranks(w_e2e), max_over_ranks(w_ie2e)
    total_in = sum_over_ranks(float(nbytes))
    total_out = sum_over_ranks(float(clen[0]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
    except Exception:
        pass
    tr_c = traffic.get("compress", {}).get("dram_per_algorithmic_byte")
    tr_i = traffic.get("inflate", {}).get("dram_per_algorithmic_byte")
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    if rank == 0:
        value = total_in * steps / t_dev_m / 1e9
        alg_bytes = (nbytes + clen[0]) * steps         # per rank: payload read + stream written
        ach = alg_bytes / t_dev / 1e9
        ach_i = alg_bytes / t_idev / 1e9
        line = {
            "metric": "BGZF compress GB/s (uncompressed)", "value": value, "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": t_dev_m / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config, "impl": "b200", "roundtrip_ok": ok,
            "ratio": total_out / total_in,
            "e2e": {"value": total_in * steps / w_e2e_m / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": clen[0],
                    "api": "b200bgzf_compress_host (pinned host buffers)"},
            "inflate": {"value": total_in * steps / t_idev_m / 1e9, "unit": "GB/s", "ms_per_step": t_idev_m / steps * 1e3,
                        "e2e": {"value": total_in * steps / w_ie2e_m / 1e9, "unit": "GB/s", "h2d_bytes_per_step": clen[0], "d2h_bytes_per_step": nbytes,
                                "api": "b200bgzf_inflate_host (pinned host buffers)"},
                        "roofline": {"bound": "hbm", "achieved": ach_i, "peak": peak, "unit": "GB/s", "frac": ach_i / peak,
                                     "traffic": int(tr_i * (nbytes + clen[0])) if tr_i else None,
                                     "traffic_source": "profiles/r01_traffic.json (ncu dram bytes per algorithmic byte) x this launch's algorithmic bytes",
                                     "kernel": "bgzf_inflate_kernel (+ member index kernels, 0.6% of the step)"}},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": int(tr_c * (nbytes + clen[0])) if tr_c else None,
                         "traffic_source": "profiles/r01_traffic.json (ncu dram bytes per algorithmic byte) x this launch's algorithmic bytes",
                         "kernel": "bgzf_compress_kernel (+ scan/gather compaction, <1% of the step)", "peak_source": peak_src,
                         "algorithmic_bytes_per_step": nbytes + clen[0]},
            "gpu_launches": launches_c + launches_i,
            "clocks": clk.summary(),
        }
        if not args.no_cpu_baseline:
            try:
                cb = cpu_reference(args, h_in.data_ptr(), nbytes, args.kind)
                if cb:
                    line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                    line["cpu_baseline"]["inflate_value"] = cb["inflate_value"]
                    line["cpu_baseline"]["ratio"] = cb["ratio"]
            except Exception as e:  # the baseline must never sink the measurement
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    codec.close()


if __name__ == "__main__":
    main()
