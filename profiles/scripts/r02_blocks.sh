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
run ad4cl_b200/libad4cl_cuda.so U768 --recorder uniform --lockstep on --local-size 768 --workload catch_at_age --n 10000000
run variants/b896.so U896 --recorder uniform --lockstep on --local-size 896 --workload catch_at_age --n 10000000
run variants/b896.so U448 --recorder uniform --lockstep on --local-size 448 --workload catch_at_age --n 10000000
run variants/b1024.so U1024 --recorder uniform --lockstep on --local-size 1024 --workload catch_at_age --n 10000000
run variants/b1024.so U512 --recorder uniform --lockstep on --local-size 512 --workload catch_at_age --n 10000000
run variants/b896.so TS896-256 --recorder uniform --lockstep off --local-size 128 --workload tape_stress --n 10000000 --tape-steps 333
run variants/b1024.so TS1024-256 --recorder uniform --lockstep off --local-size 256 --workload tape_stress --n 10000000 --tape-steps 333
run ad4cl_b200/libad4cl_cuda.so TS768-256 --recorder uniform --lockstep off --local-size 256 --workload tape_stress --n 10000000 --tape-steps 333
