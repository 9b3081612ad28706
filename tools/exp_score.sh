for mb in 4 5; do
  PB200_NVCC_FLAGS="-DPB_SCORE_MIN_BLOCKS=$mb" python -m porechop_b200.build --force > /dev/null
  echo "== score minb $mb"
  for i in 1 2; do timeout 900 python bench.py --workload middle --steps 3 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['gcups'],1), round(d['ms_per_step'],3))"; done
done
