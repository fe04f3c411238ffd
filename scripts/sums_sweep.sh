# shape / chunk sweep of ocb_budget_sums_kernel at the bench workload (1024^3 Float64, 7 integrals)
run() { OCB_SUMS_CONFIG=$1 OCB_FUSED_KCHUNK=$2 timeout 300 python bench.py --steps ${3:-10} --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('cfg=$1 kchunk=$2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('cfg=$1 kchunk=$2 ERR', e)
"; }
for cfg in ${CFGS:-"" 8,4,3,1 4,6,3,1 12,2,3,1 8,3,3,1 4,5,3,2 4,4,4,2}; do
  for kc in ${KCS:-""}; do run "$cfg" "$kc" ${STEPS:-10}; done
done
