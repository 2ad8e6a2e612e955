mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r02i_tests.log 2>&1
tail -8 gpurun_out/r02i_tests.log
run() { # lib, label, bench args...
  lib=$1; label=$2; shift 2
  AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" 2>/dev/null | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$lib', '$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], d['launch'])
except Exception as e:
    print('$lib', '$label', 'FAILED', repr(e))
"
}
for lib in ad4cl_b200/libad4cl_cuda.so variants/lockstep.so; do
  run $lib caa-uniform --recorder uniform --workload catch_at_age --n 10000000
  run $lib caa-lane-pf0 --recorder lane --prefetch off --workload catch_at_age --n 10000000
  run $lib caa-lane-pf1 --recorder lane --prefetch on --workload catch_at_age --n 10000000
  run $lib caa-lane-pf1-w32 --recorder lane --prefetch on --window 32 --workload catch_at_age --n 10000000
  run $lib caa-uniform-w32 --recorder uniform --window 32 --workload catch_at_age --n 10000000
  run $lib caa-auto --workload catch_at_age --n 10000000
  run $lib ts-uniform --recorder uniform --workload tape_stress --n 10000000 --tape-steps 333
  run $lib ts-auto --workload tape_stress --n 10000000 --tape-steps 333
done 2>&1 | tee gpurun_out/r02i_ab.log
