set -x
python -m pytest tests/test_fused_budget_gpu.py -x -q -k "not full_size" > gpurun_out/s3_fused_tests.log 2>&1; tail -15 gpurun_out/s3_fused_tests.log
for cfg in "" 8,3,4,1 8,4,3,1 8,2,3,2 8,2,4,2 12,2,3,1 16,2,3,1 4,4,3,2 4,3,4,3 8,3,2,2; do
  for kc in "" 256 128; do
    echo "== cfg=$cfg kchunk=$kc"
    OCB_SUMS_CONFIG=$cfg OCB_FUSED_KCHUNK=$kc timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
except Exception as e: print('ERR', e)
"
  done
done > gpurun_out/s3_sums_sweep.log 2>&1
cat gpurun_out/s3_sums_sweep.log
