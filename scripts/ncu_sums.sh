# ncu evidence for the bench command (class-sum kernel): launch list with DRAM traffic of every library kernel in the timed
# region, one --set full capture of the sweep kernel; plus one --set full capture of the field-writing variant (--fields)
set -x
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sums_kernel|finalize" -c 12 --csv --log-file gpurun_out/r1_launches_bench1024_sums_v2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sums_kernel -c 1 -o gpurun_out/r1_sums_v2 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_ncu_full.log 2>&1
python bench.py --fields --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_plain_fields.log 2>&1 || exit 1
tail -1 gpurun_out/s3_plain_fields.log | python scripts/benchline.py fields
ncu --set full --clock-control none --import-source on -k regex:ocb_budget_kernel -c 1 -o gpurun_out/r1_budget_fields_v13 -f python bench.py --fields --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s3_ncu_full_fields.log 2>&1
