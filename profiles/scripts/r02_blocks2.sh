mkdir -p gpurun_out
run() { # lib label, bench args...
  lib=$1; label=$2; shift 2
  AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/err.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$lib', '$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], l['grid'], l['block'], l['smem_bytes'], l['recorder'], 'pf', l['sweep_prefetch'], 'ls', l['lockstep'])
except Exception as e:
    print('$lib', '$label', 'FAILED', repr(e)); print(open('gpurun_out/err.log').read()[-400:])
"
}
run variants/b640.so U640 --recorder uniform --lockstep on --local-size 640 --workload catch_at_age --n 10000000
run variants/b640.so U320 --recorder uniform --lockstep on --local-size 320 --workload catch_at_age --n 10000000
run variants/b512.so U512 --recorder uniform --lockstep on --local-size 512 --workload catch_at_age --n 10000000
run variants/b512.so U256 --recorder uniform --lockstep on --local-size 256 --workload catch_at_age --n 10000000
run variants/b640.so TS640-320 --recorder uniform --lockstep off --local-size 320 --workload tape_stress --n 10000000 --tape-steps 333
run variants/b512.so TS512-256 --recorder uniform --lockstep off --local-size 256 --workload tape_stress --n 10000000 --tape-steps 333
