#!/bin/bash
# A/B of library builds on ONE box in ONE gpurun call (box-to-box spread is ~1 %, the same order as
# most of the steps in profiles/README.md).  Build variants out of tree, e.g.
#   make -C ad4cl_b200/csrc BUILD=/tmp/b1 TARGET=$PWD/variants/v1.so EXTRA=-DAD4CL_SWEEP_UNROLL=8
# and run  gpurun -- 'profiles/ab.sh variants/base.so variants/v1.so'
# (variants/ is git-ignored but travels to the GPU box; host.py loads $AD4CL_CUDA_LIB).
for rep in 1 2; do
for lib in "$@"; do
  for wl in "catch_at_age --n 10000000" "tape_stress --n 10000000 --tape-steps 333" "regression_linear --n 1000000" \
            "regression_logistic --n 1000000" "linreg2_simple --n 1000000"; do
    AD4CL_CUDA_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', '$wl'.split()[0], 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'],
      'frac=%.3f' % d['roofline']['frac'], 'prefetch=' + d['config']['sweep_prefetch'].split()[0])
"
  done
done
done
