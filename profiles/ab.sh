#!/bin/bash
# A/B of library builds on ONE box in ONE gpurun call (box-to-box spread is ~1 %, the same order as
# most of the steps in profiles/README.md).  Build variants out of tree, e.g.
#   make -C ad4cl_b200/csrc BUILD=/tmp/b1 TARGET=$PWD/variants/v1.so EXTRA=-DAD4CL_SOME_SWITCH=1
# and run  gpurun -- 'profiles/ab.sh ad4cl_b200/libad4cl_cuda.so variants/v1.so'
# (variants/ is git-ignored but travels to the GPU box; host.py loads $AD4CL_CUDA_LIB).
REPS=${REPS:-2}
for rep in $(seq 1 $REPS); do
for lib in "$@"; do
 for rec in ${RECS:-auto}; do
  for wl in "catch_at_age --n 10000000" "tape_stress --n 10000000 --tape-steps 333" "regression_linear --n 1000000" \
            "regression_logistic --n 1000000" "linreg2_simple --n 1000000"; do
    AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --recorder ${rec%%+*} $( [ "${rec##*+}" = "pf" ] && echo --prefetch on || ( [ "$rec" = "lane" ] && echo --prefetch off ) ) --workload $wl 2>/dev/null | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$lib', '$wl'.split()[0], 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'ms_per_step=%.4f' % d['ms_per_step'],
          'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], 'frac=%.3f' % d['roofline']['frac'], 'tier=' + d['launch']['tape_tier'], 'rec=$rec->' + str(d['launch']['recorder']) + '/pf' + str(d['launch']['sweep_prefetch']))
except Exception as e:
    print('$lib', '$wl'.split()[0], 'FAILED', repr(e))
"
  done
 done
done
done
