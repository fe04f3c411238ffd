# filter tests, timing of config 5 (17-tap Gaussian, 512x512x256 Float64) and the ncu launch list of one staged filter
python -m pytest tests/test_filters_gpu.py tests/test_filtered_ke_gpu.py -x -q 2>&1 | tail -3
for env in "" "OCB_FILTER_NO_SIDE_STREAMS=1"; do
  echo "== $env"
  env $env python scripts/bench_kernels.py filter 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    try:
        d = json.loads(line)
        print(d['ms'], d['name'][:100])
    except Exception:
        pass
"
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ocb_filter --launch-skip 9 -c 9 --csv --log-file gpurun_out/r1_launches_filter_v3.csv python scripts/bench_kernels.py filter_only > gpurun_out/s3_filter_ncu.log 2>&1
