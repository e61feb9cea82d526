#!/bin/bash
# One-GPU evidence run (under gpurun): bench lines of every workload, ncu captures of the two
# tracking kernels, the launch list of the default bench.  Writes gpurun_out/r02/.
set -u
O=gpurun_out/r02; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.txt
python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench_n1.err
: > $O/bench_workloads.jsonl
for w in default_yaml absdom thick; do python bench.py --workload $w --steps 3 --warmup 3 >> $O/bench_workloads.jsonl 2>> $O/bench_workloads.err; done
python bench.py --workload hetero_1e6 --steps 1 --warmup 1 >> $O/bench_workloads.jsonl 2>> $O/bench_workloads.err
python bench.py --workload single --rng philox --steps 3 --warmup 3 --no-cpu-baseline >> $O/bench_workloads.jsonl 2>> $O/bench_workloads.err
python bench.py --workload culayer --steps 2 >> $O/bench_workloads.jsonl 2>> $O/bench_workloads.err
# launch list of the default bench (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bench_n1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench_n1.out 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 1 -c 1 -f -o $O/track_kernel \
    python tools/probe.py 2e7 > $O/ncu_track.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:world_kernel -s 1 -c 1 -f -o $O/world_kernel_thin125 \
    python tools/probe_thin.py 125 112000000 > $O/ncu_world_thin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:world_kernel -s 1 -c 1 -f -o $O/world_kernel_1000 \
    python tools/probe_thin.py 1000 14000000 > $O/ncu_world_1000.log 2>&1
ls -la $O
