for env in "" "OCB_FILTER_NO_PAIRS=1"; do
  tag=$( [ -z "$env" ] && echo pairs || echo nopairs )
  env $env ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ocb_filter --launch-skip 9 -c 9 --csv --log-file gpurun_out/s3_filter_launches_$tag.csv python scripts/bench_kernels.py filter_only > gpurun_out/s3_filter_ncu_$tag.log 2>&1
done
