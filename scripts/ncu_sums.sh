# one ncu --set full capture of the class-sum kernel at the bench workload + the launch list of the bench command
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sums_kernel -c 1 -o gpurun_out/r1_sums_v1 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ocb_ -c 12 --csv --log-file gpurun_out/r1_launches_bench1024_sums_v1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_ncu_list.log 2>&1
