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
L=variants/upf.so
run $L ts-pf0 --recorder uniform --prefetch off --lockstep off --local-size 256 --workload tape_stress --n 10000000 --tape-steps 333
run $L ts-pf1 --recorder uniform --prefetch on --lockstep off --local-size 256 --workload tape_stress --n 10000000 --tape-steps 333
run $L ts-auto --workload tape_stress --n 10000000 --tape-steps 333
run $L tsL-pf0 --recorder uniform --prefetch off --lockstep off --local-size 256 --workload tape_stress --n 2000000 --tape-steps 3333
run $L tsL-pf1 --recorder uniform --prefetch on --lockstep off --local-size 256 --workload tape_stress --n 2000000 --tape-steps 3333
run $L ts100-pf0 --recorder uniform --prefetch off --lockstep off --local-size 256 --workload tape_stress --n 10000000 --tape-steps 33
run $L ts100-pf1 --recorder uniform --prefetch on --lockstep off --local-size 256 --workload tape_stress --n 10000000 --tape-steps 33
run $L caa-U768-pf0 --recorder uniform --prefetch off --lockstep on --local-size 768 --workload catch_at_age --n 10000000
run $L caa-U768-pf1 --recorder uniform --prefetch on --lockstep on --local-size 768 --workload catch_at_age --n 10000000
run $L caa-auto --workload catch_at_age --n 10000000
AD4CL_CUDA_LIB=$PWD/$L timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -x 2>&1 | tail -4
