#!/bin/bash
# Round-2 profile capture of the configuration chosen from tools/r2_first_call.sh (one GPU, never under torchrun):
#   gpurun --timeout 1500 -- 'bash tools/r2_ncu.sh "--opt profile=1 --opt short2p=1 --opt tight_window=1"'
# 1) launch list (kernel shares of a step) of the bench command, 2) `--set full` of the DP kernels, both brought back in
# gpurun_out/r2_ncu/ ; summarise here with tools/ncu_summary.py and copy the summaries into profiles/.
set -u
opts="${1:-}"
out=gpurun_out/r2_ncu
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline $opts > $out/bench_under_ncu.log 2>&1
for kern in trace_kernel score_kernel; do
    wl=""; [ $kern = score_kernel ] && wl="--workload middle --reads 65536"
    ncu --set full --clock-control none --import-source on -k regex:$kern -s 4 -c 4 -o $out/$kern -f \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline $wl $opts > $out/${kern}_under_ncu.log 2>&1
done
ls -la $out
