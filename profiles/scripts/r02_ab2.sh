mkdir -p gpurun_out
run() { # lib label, bench args...
  lib=$1; label=$2; shift 2
  AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/err.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$lib', '$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], l['grid'], l['block'], l['smem_bytes'], l['tape_tier'], l['recorder'], 'pf', l['sweep_prefetch'], 'ls', l['lockstep'])
except Exception as e:
    print('$label', 'FAILED', repr(e)); print(open('gpurun_out/err.log').read()[-600:])
"
}
for rep in 1 2; do
for lib in ad4cl_b200/libad4cl_cuda.so $(ls variants/*.so); do
run $lib caa-auto --workload catch_at_age --n 10000000
run $lib caa-U256 --recorder uniform --lockstep on --local-size 256 --workload catch_at_age --n 10000000
run $lib ts-auto --workload tape_stress --n 10000000 --tape-steps 333
done
done 2>&1 | tee gpurun_out/r02_ab2.log
