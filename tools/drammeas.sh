timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for mb in 128 72; do
  echo "== scratch_mb $mb"
  PB200_SCRATCH_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['launch_ms'], d['gpu_launches'])"
  PB200_SCRATCH_MB=$mb timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:trace_kernel -s 2 -c 1 --csv --log-file gpurun_out/dram_s$mb.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/dram_s$mb.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
print([ (r[-3], r[-1]) for r in rows[hi+1:]])
PY
done
