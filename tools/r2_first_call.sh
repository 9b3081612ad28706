#!/bin/bash
# Round-2 first gpurun call: GPU parity of the opt-in paths prepared at the end of round 1 (no GPU was left to run them;
# they pass in the host simulation, tests/test_sim_engine.py), then one bench line per option so that the defaults can be
# decided from measurements.
#   gpurun --timeout 3000 -- 'bash tools/r2_first_call.sh'          (everything: about 45 GPU-minutes)
#   gpurun --timeout 1500 -- 'R2_STAGES="tests bench1" bash tools/r2_first_call.sh'     (stages: tests bench1 bench2 sanitize v2)
# Everything lands in gpurun_out/r2_first/ ; summary.txt is the file to read.
set -u
STAGES="${R2_STAGES:-tests bench1 bench2 sanitize v2}"
stage() { case " $STAGES " in *" $1 "*) return 0;; *) return 1;; esac; }
out=gpurun_out/r2_first
mkdir -p $out
if stage tests; then
python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_zz_options.py > $out/pytest_gpu_default.log 2>&1
echo "gpu tests (default paths) rc=$?" | tee -a $out/summary.txt
timeout 900 python -m pytest tests/test_gpu_zz_options.py -m gpu -q > $out/pytest_gpu_options.log 2>&1
echo "gpu tests (opt-in paths) rc=$?" | tee -a $out/summary.txt
tail -n 3 $out/pytest_gpu_default.log $out/pytest_gpu_options.log >> $out/summary.txt
fi
run() {   # name, bench args...
    name=$1; shift
    timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" > $out/bench_$name.json 2> $out/bench_$name.err
    python - "$out/bench_$name.json" "$name" <<'PY' | tee -a $out/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('%-30s value %.3e  e2e %.3e  dp launch %.3f ms x %.1f/step  e2e step %.2f ms' % (
        sys.argv[2], d['value'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['launches_per_step'],
        d['e2e']['ms_per_step']))
except Exception as e:
    print('%-30s FAILED %r' % (sys.argv[2], e))
PY
}
if stage bench1; then
# config 1 (headline): e2e levers, kernel levers, everything
run endtrim_default
run endtrim_multi --e2e-multi
run endtrim_pack_multi --opt h2d_pack=1 --e2e-multi
run endtrim_profile --opt profile=1
run endtrim_profile_scratch160 --opt profile=1 --opt scratch_mb=160      # 85 registers allow 6 blocks/SM; the 128 MB scratch cap holds it at 5
run endtrim_short2p_tight --opt short2p=1 --opt tight_window=1
run endtrim_profile_short2p_tight --opt profile=1 --opt short2p=1 --opt tight_window=1
run endtrim_all --opt profile=1 --opt short2p=1 --opt tight_window=1 --opt h2d_pack=1 --e2e-multi
run endtrim_all_decisions --opt profile=1 --opt short2p=1 --opt tight_window=1 --opt h2d_pack=1 --e2e-decisions
fi
if stage bench2; then
# config 3 (demux) and config 4 (middle scan)
run demux_default --workload demux
run demux_short2p_tight --workload demux --opt short2p=1 --opt tight_window=1
run demux_profile --workload demux --opt profile=1
run demux_profile_short2p_tight --workload demux --opt profile=1 --opt short2p=1 --opt tight_window=1
run demux_decisions --workload demux --e2e-decisions
run middle_default --workload middle
run middle_profile_tight --workload middle --opt profile=1 --opt tight_window=1
run middle_profile_tight_pack --workload middle --opt profile=1 --opt tight_window=1 --opt h2d_pack=1
cat $out/summary.txt
# config 5 (read-length sweep x 192 barcodes), aggregate over the lengths
run sweep_default --workload sweep
run sweep_profile_tight --workload sweep --opt profile=1 --opt tight_window=1
cat $out/summary.txt
fi
if stage sanitize; then
# hardware-side sanitizers on tiny inputs (every path once)
for tool in memcheck racecheck; do
    timeout 480 compute-sanitizer --tool $tool python tools/r2_sanitize.py > $out/sanitize_$tool.log 2>&1
    echo "compute-sanitizer $tool rc=$? : $(grep -c 'ERROR SUMMARY' $out/sanitize_$tool.log) summaries, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/sanitize_$tool.log | tail -n 1)" | tee -a $out/summary.txt
done
fi
if stage v2; then
# compile-time experiment: single-step traceback (rebuilds the library; keep this last)
PB200_NVCC_FLAGS=-DPB_TRACEBACK_V2 python -m porechop_b200.build --force > $out/build_v2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_gpu_v2.log 2>&1; echo "gpu parity with PB_TRACEBACK_V2 rc=$?" | tee -a $out/summary.txt
run endtrim_tracebackv2
run endtrim_tracebackv2_all --opt profile=1 --opt short2p=1 --opt tight_window=1
python -m porechop_b200.build --force > /dev/null 2>&1
fi
cat $out/summary.txt
