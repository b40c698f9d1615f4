#!/bin/bash
# One GPU-box call that produces every measured artefact of a round under gpurun_out/ (then summarise here:
# tools/launch_summary.py, tools/ncu_summary.py, tools/ncu_funcs.py, tools/sass_summary.py, tools/traffic_from_ncu.py).
#   gpurun --timeout 1800 -- 'bash tools/capture_round.sh r02'
R=${1:-r02}; O=gpurun_out; mkdir -p $O
python bench.py --steps 5 --warmup 3 > $O/bench_$R.json 2> $O/bench_$R.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${R}_reference.json 2>> $O/bench_$R.err
python bench.py --level 12 --mib 1024 --steps 1 --warmup 3 > $O/bench_${R}_L12.json 2>> $O/bench_$R.err
python bench.py --workload config5 > $O/bench_${R}_config5.json 2>> $O/bench_$R.err
python tools/level_sweep.py 256 > $O/levels_$R.jsonl 2>> $O/bench_$R.err
(for l in 1 6 9 12; do python tools/phase_profile.py $l $([ $l = 12 ] && echo 16 || echo 64) fastq; done; python tools/phase_profile.py 6 64 sam) > $O/phases_$R.txt 2>&1
python tools/pcie_probe.py > $O/pcie_$R.txt 2>&1
python tools/e2e_probe.py 1024 >> $O/pcie_$R.txt 2>&1
bash tools/applet_compare.sh 1024 > $O/applet_$R.txt 2>&1
python tools/container_bench.py 1024 256 > $O/containers_$R.jsonl 2>> $O/bench_$R.err
(python tools/compress_fuzz.py 400 7; python tools/compress_fuzz.py 300 23) > $O/compress_fuzz_$R.txt 2>&1
(timeout 300 python tools/container_fuzz.py 300 5; timeout 200 python tools/container_fuzz.py 150 9) > $O/container_fuzz_$R.txt 2>&1
# launch list of the bench command (shares per kernel; times under ncu are cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$R.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_bench_$R.log 2>&1
# one full capture of each dominant kernel
ncu --set full --clock-control none --import-source on -k regex:bgzf_compress_kernel --launch-skip 1 --launch-count 1 -f -o $O/prof_compress_$R python tools/prof_run.py compress 148 6 fastq 2 > $O/ncu_compress_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate_kernel --launch-skip 1 --launch-count 1 -f -o $O/prof_inflate_$R python tools/prof_run.py inflate 1024 6 fastq 2 > $O/ncu_inflate_$R.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > $O/gpu_$R.txt
ls -la $O | tail -20
