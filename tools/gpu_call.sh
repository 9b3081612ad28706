#!/bin/bash
# One gpurun call of round 2 in stages (everything lands in gpurun_out/<tag>/, summary.txt is the file to read):
#   gpurun --timeout 2700 -- 'STAGES="tests sanitize bench ab ncu" TAG=r2b bash tools/gpu_call.sh'
#   tests     smoke + the whole GPU tier (incl. the reference's own 85 CLI tests on the CUDA engine)
#   sanitize  compute-sanitizer memcheck + racecheck over every engine path (tools/sanitize_paths.py)
#   bench     the driver's bench line (N = 1, all configs) + the reference arm
#   ab        run-time A/B of the headline config: packed upload, chunk sizes (compile-time A/B: tools/gpu_ab.sh)
#   ncu       launch list of the bench command + `--set full` captures of the DP kernels (never a bench value)
set -u
STAGES="${STAGES:-tests bench}"
TAG="${TAG:-r2}"
stage() { case " $STAGES " in *" $1 "*) return 0;; *) return 1;; esac; }
out=gpurun_out/$TAG
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/gpu.txt 2>&1
lscpu | head -n 25 > $out/host.txt 2>&1; nproc >> $out/host.txt; cat /sys/fs/cgroup/cpu.max >> $out/host.txt 2>/dev/null; free -g >> $out/host.txt
line() {   # file, label
    python - "$1" "$2" <<'PY' | tee -a $out/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline') or {}
    a = d.get('roofline_alu') or {}
    print('%-26s value %.3e  e2e %.3e (%.2f ms, x%.2f of value)  dom %s %.3f ms x %.1f  alu frac %s  parity %s' % (
        sys.argv[2], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['value'] / d['value'], r.get('kernel'),
        r.get('launch_ms') or 0, r.get('launches_per_step') or 0, a.get('frac'), d.get('parity')))
    for k, v in (d.get('configs') or {}).items():
        if 'value' in v:
            print('   %-23s value %.3e  e2e %.3e (x%.2f)  gcups %.0f  parity %s  cpu %s' % (
                k, v['value'], v['e2e']['value'], v['e2e']['value'] / v['value'], v.get('gcups', 0), v.get('parity'),
                (v.get('cpu_baseline') or {}).get('value')))
        else:
            print('   %-23s %s' % (k, str(v)[:300]))
    if 'cpu_baseline' in d:
        c = d['cpu_baseline']
        print('   cpu: best %.3e reads/s (%s), 1 thread %.3e; sweep %s' % (c['value'], c.get('mode'), c['one_thread']['value'],
              [(r['threads'], r['procs'], round(r['reads_per_s'])) for r in c.get('sweep', [])]))
except Exception as e:
    print('%-26s FAILED %r' % (sys.argv[2], e))
PY
}
if stage tests; then
python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1
echo "gpu tests rc=$?" | tee -a $out/summary.txt
tail -n 4 $out/pytest_gpu.log >> $out/summary.txt
fi
if stage sanitize; then
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool python tools/sanitize_paths.py > $out/sanitize_$tool.log 2>&1
    echo "compute-sanitizer $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/sanitize_$tool.log | tail -n 1) ; $(grep -c 'SANITIZE RUN COMPLETE' $out/sanitize_$tool.log) complete" | tee -a $out/summary.txt
done
fi
if stage bench; then
timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?" | tee -a $out/summary.txt
line $out/bench_n1.json bench_n1
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference arm rc=$?" | tee -a $out/summary.txt
python - $out/bench_reference.json <<'PY' | tee -a $out/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('reference arm: value %.3e (%s), 1 thread %.3e; configs %s' % (d['value'], d['cpu_baseline']['mode'], d['cpu_baseline']['one_thread']['value'],
          {k: (round(v['value'], 1) if 'value' in v else v) for k, v in d['configs'].items()}))
except Exception as e:
    print('reference arm FAILED %r' % e)
PY
fi
run() { name=$1; shift; timeout 400 python bench.py --steps 8 --warmup 3 --configs none --no-cpu-baseline "$@" > $out/bench_$name.json 2> $out/bench_$name.err; line $out/bench_$name.json $name; }
if stage ab; then
run endtrim_default
run endtrim_pack --opt h2d_pack=1
run endtrim_pack64 --opt h2d_pack=1 --opt pack_threads=64
run endtrim_chunk64k --opt chunk_tasks=65536
run endtrim_chunk256k --opt chunk_tasks=262144
run endtrim_pack_chunk256k --opt h2d_pack=1 --opt chunk_tasks=262144
run middle_default --workload middle --config-reads middle=262144
run middle_pack --workload middle --opt h2d_pack=1
python -m porechop_b200.build --force > /dev/null 2>&1
run demux_bytes --workload demux --reads 65536
fi
if stage ncu; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_endtrim.csv \
    python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 2 -o $out/trace_kernel -f \
    python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline > $out/trace_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 1 -c 1 -o $out/score_kernel -f \
    python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline --workload middle --reads 262144 > $out/score_under_ncu.log 2>&1
ls -la $out | tee -a $out/summary.txt
fi
cat $out/summary.txt
