mkdir -p gpurun_out
run() { # lib label, bench args...
  lib=$1; label=$2; shift 2
  AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/err.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$lib', '$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], l['grid'], l['block'], l['recorder'], 'pf', l['sweep_prefetch'], 'ls', l['lockstep'], d.get('parity_check', {}).get('status'))
except Exception as e:
    print('$lib', '$label', 'FAILED', repr(e)); print(open('gpurun_out/err.log').read()[-400:])
"
}
for L in ad4cl_b200/libad4cl_cuda.so variants/snake.so; do
run $L caa-auto-10M --workload catch_at_age --n 10000000
run $L caa-auto-1.25M --workload catch_at_age --n 1250000
run $L caa-U768-1.25M --recorder uniform --lockstep on --local-size 768 --workload catch_at_age --n 1250000
run $L caa-L256-1.25M --recorder lane --prefetch on --lockstep off --local-size 256 --workload catch_at_age --n 1250000
run $L caa-L256-100k --recorder lane --prefetch on --lockstep off --local-size 256 --workload catch_at_age --n 113600
run $L ts-auto --workload tape_stress --n 10000000 --tape-steps 333
run $L linreg2 --workload linreg2_simple --n 1000000
done
AD4CL_CUDA_LIB=$PWD/variants/snake.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_divergence.py tests/test_gpu_ops.py -m gpu -q --tb=short -x 2>&1 | tail -4
