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
for rep in 1 2; do
for ls in 768 384 256 192; do run ad4cl_b200/libad4cl_cuda.so U$ls --recorder uniform --lockstep on --local-size $ls --workload catch_at_age --n 10000000; done
done
for ls in 384 768; do run ad4cl_b200/libad4cl_cuda.so U$ls-1.25M --recorder uniform --lockstep on --local-size $ls --workload catch_at_age --n 1250000; done
for ls in 384; do run ad4cl_b200/libad4cl_cuda.so L$ls --recorder lane --prefetch on --lockstep on --local-size $ls --workload catch_at_age --n 10000000; done
for ls in 384; do run ad4cl_b200/libad4cl_cuda.so TS$ls --recorder uniform --lockstep off --local-size $ls --workload tape_stress --n 10000000 --tape-steps 333; done
