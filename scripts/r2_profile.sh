#!/bin/bash
# Round-2 measurement pass on one B200: GPU tests, the default bench line, an A/B of the ring-mode row classes, and the ncu
# captures bench.py's roofline.traffic is parsed from (run AFTER the plain bench has exited 0).
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_gputest2.log
tail -3 $O/r2_gputest2.log
python bench.py --steps 20 --warmup 3 > $O/r2_bench_n1_b.json 2> $O/r2_bench_n1_b.err; echo "bench rc=$?"
OCB_SUMS_NO_RING=1 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2_bench_n1_noring.json 2>> $O/r2_bench_n1_b.err
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/r2_bench_n1_ring.json 2>> $O/r2_bench_n1_b.err
python - <<'PY'
import json
for f in ("r2_bench_n1_b", "r2_bench_n1_noring", "r2_bench_n1_ring"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["roofline"]["frac"], d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
NCU="ncu --clock-control none --csv"
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:ocb_budget -c 8 --log-file $O/r2_traffic_bench1024_sums.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/ncu1.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:ocb_budget -c 8 --log-file $O/r2_traffic_bench1024_fields.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --fields > $O/ncu2.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --log-file $O/r2_launches_bench1024.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-secondary > $O/ncu3.log 2>&1
tail -2 $O/r2_traffic_bench1024_sums.csv | cut -c1-300
