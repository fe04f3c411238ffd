# round-end validation on one B200: the whole GPU suite, smoke(), the default bench line and the reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log | python scripts/benchline.py default
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.log 2>&1; tail -1 gpurun_out/final_bench_reference.log | cut -c1-600
