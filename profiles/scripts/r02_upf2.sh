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
for L in ad4cl_b200/libad4cl_cuda.so variants/upf.so; do
run $L caa-auto --workload catch_at_age --n 10000000
run $L ts-auto --workload tape_stress --n 10000000 --tape-steps 333
done
done
L=variants/upf.so
run $L tsL-auto --workload tape_stress --n 2000000 --tape-steps 3333
run $L ts100-auto --workload tape_stress --n 10000000 --tape-steps 33
