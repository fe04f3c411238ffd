#!/bin/bash
# quick GPU pass: the tests named on the command line, then the kernel timings named in $KB (bench_kernels.py arguments);
# $AB = an environment assignment to A/B the `filter_only` timing with (e.g. OCB_FILTER_NO_XY=1)
O=gpurun_out
python -m pytest $@ -x -q 2>&1 | tail -12 > $O/r2_quick_tests.log; tail -4 $O/r2_quick_tests.log
python scripts/bench_kernels.py ${KB:-bpe} > $O/r2_quick_kernels.log 2>&1; cat $O/r2_quick_kernels.log | cut -c1-250
if [ -n "${AB:-}" ]; then
  echo "A/B with $AB:"; env $AB python scripts/bench_kernels.py filter_only 2>&1 | cut -c1-200; echo "without:"; python scripts/bench_kernels.py filter_only 2>&1 | cut -c1-200
fi
