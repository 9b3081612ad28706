#!/bin/bash
# Round-2 first gpurun call: parity of the opt-in paths prepared at the end of round 1 (no GPU was left to run them),
# then one bench line per option so the defaults can be decided from measurements.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# Everything lands in gpurun_out/r2_first/.
set -u
out=gpurun_out/r2_first
mkdir -p $out
python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.txt
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_zz_options.py > $out/pytest_gpu_default.log 2>&1; echo "gpu tests (default paths) rc=$?" | tee -a $out/summary.txt
python -m pytest tests/test_gpu_zz_options.py -m gpu -q > $out/pytest_gpu_options.log 2>&1; echo "gpu tests (opt-in paths) rc=$?" | tee -a $out/summary.txt
tail -n 3 $out/pytest_gpu_default.log $out/pytest_gpu_options.log >> $out/summary.txt
run() {   # name, bench args...
    name=$1; shift
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $out/bench_$name.json 2> $out/bench_$name.err
    python - "$out/bench_$name.json" "$name" <<'PY' | tee -a $out/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('%-28s value %.3e  e2e %.3e  dp launch %.3f ms x %.1f/step' % (sys.argv[2], d['value'], d['e2e']['value'],
          d['roofline']['launch_ms'], d['roofline']['launches_per_step']))
except Exception as e:
    print('%-28s FAILED %r' % (sys.argv[2], e))
PY
}
run endtrim_default
run endtrim_multi --e2e-multi
run endtrim_pack --opt h2d_pack=1
run endtrim_pack_multi --opt h2d_pack=1 --e2e-multi
run endtrim_pack16_multi --opt h2d_pack=1 --opt pack_threads=16 --e2e-multi
run endtrim_short2p --opt short2p=1
run endtrim_short2p_tight --opt short2p=1 --opt tight_window=1
run endtrim_profile --opt profile=1
run endtrim_profile_short2p_tight --opt profile=1 --opt short2p=1 --opt tight_window=1
run endtrim_all --opt profile=1 --opt short2p=1 --opt tight_window=1 --opt h2d_pack=1 --e2e-multi
run endtrim_decisions --e2e-decisions
run endtrim_decisions_pack --e2e-decisions --opt h2d_pack=1
run demux_default --workload demux
run demux_decisions --workload demux --e2e-decisions
run demux_short2p_tight --workload demux --opt short2p=1 --opt tight_window=1
run middle_default --workload middle
run middle_tight --workload middle --opt tight_window=1
run middle_profile --workload middle --opt profile=1
run middle_profile_rowoff --workload middle --opt profile=1 --opt rowoff=1
run middle_profile_tight_pack --workload middle --opt profile=1 --opt tight_window=1 --opt h2d_pack=1
run middle_tight_pack --workload middle --opt tight_window=1 --opt h2d_pack=1
cat $out/summary.txt
