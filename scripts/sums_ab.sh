python -m pytest tests/test_fused_budget_gpu.py tests/test_multigpu_gpu.py -x -q -k "not full_size" 2>&1 | tail -3
for rep in 1 2; do
  for env in "OCB_SUMS_WIDE_TILES=1" "X=1" "OCB_SUMS_CONFIG=4,6,3,1"; do
    env $env python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python scripts/benchline.py "$env"
  done
done
