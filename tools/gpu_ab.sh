#!/bin/bash
# Compile-time A/B on the GPU box: every VARIANT = "name|nvcc flags|bench.py arguments" rebuilds cpp_functions.so with the
# flags (PB200_NVCC_FLAGS), optionally runs the GPU parity tests (PARITY=1) and prints the bench line's key numbers.
#   gpurun --timeout 1500 -- 'TAG=r2d PARITY=1 bash tools/gpu_ab.sh "base||" "or11|-DPB_FLAG_ALU_A=0x1 -DPB_FLAG_ALU_B=0x1|"'
# UBENCH=1 first builds tools/ubench_pipes.cu, runs it plain and under ncu (pipe counters of every kernel).
set -u
TAG="${TAG:-ab}"
out=gpurun_out/$TAG
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
if [ "${UBENCH:-0}" = "1" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $out/ubench_pipes tools/ubench_pipes.cu > $out/ubench_build.log 2>&1
    $out/ubench_pipes > $out/ubench_pipes.txt 2>&1
    ncu --clock-control none --csv --log-file $out/ubench_pipes_ncu.csv --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active \
        $out/ubench_pipes > $out/ubench_under_ncu.txt 2>&1
    cat $out/ubench_pipes.txt | tee -a $out/summary.txt
fi
for v in "$@"; do
    name="${v%%|*}"; rest="${v#*|}"; flags="${rest%%|*}"; bargs="${rest#*|}"
    PB200_NVCC_FLAGS="$flags" python -m porechop_b200.build --force > $out/build_$name.log 2>&1 || { echo "$name BUILD FAILED" | tee -a $out/summary.txt; continue; }
    if [ "${PARITY:-0}" = "1" ]; then
        timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_$name.log 2>&1; echo "$name parity rc=$?" | tee -a $out/summary.txt
    fi
    timeout 400 python bench.py --steps 8 --warmup 3 --configs none --no-cpu-baseline $bargs > $out/bench_$name.json 2> $out/bench_$name.err
    python - $out/bench_$name.json "$name" "$flags $bargs" <<'PY' | tee -a $out/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline') or {}
    print('%-22s value %.4e  e2e %.4e  dom %s %.3f ms x %.1f  kernels %s   [%s]' % (
        sys.argv[2], d['value'], d['e2e']['value'], r.get('kernel'), r.get('launch_ms') or 0, r.get('launches_per_step') or 0,
        {k: round(v, 3) for k, v in (r.get('kernels_ms_per_step') or {}).items()}, sys.argv[3]))
except Exception as e:
    print('%-22s FAILED %r' % (sys.argv[2], e))
PY
done
python -m porechop_b200.build --force > /dev/null 2>&1
if [ "${FLAT:-0}" = "1" ]; then      # flat FASTQ pipeline (bytes -> trimmed bytes) on this host + GPU, per-stage seconds
    timeout 600 python tools/flat_pipeline_bench.py --reads ${FLAT_READS:-200000} --repeat 3 > $out/flat_pipeline.json 2> $out/flat_pipeline.err
    echo "flat pipeline: $(cat $out/flat_pipeline.json | head -c 600)" | tee -a $out/summary.txt
fi
if [ "${NCU:-0}" = "1" ]; then       # launch list + full captures of the default build (never a bench value)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_endtrim.csv \
        python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 2 -o $out/trace_kernel -f \
        python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline > $out/trace_under_ncu.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 1 -c 1 -o $out/score_kernel -f \
        python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline --workload middle --reads 262144 > $out/score_under_ncu.log 2>&1
fi
cat $out/summary.txt
