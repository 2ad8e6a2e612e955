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
for ls in 672 224 336; do run variants/b672.so U$ls --recorder uniform --lockstep on --local-size $ls --workload catch_at_age --n 10000000; done
for ls in 704 352; do run variants/b704.so U$ls --recorder uniform --lockstep on --local-size $ls --workload catch_at_age --n 10000000; done
for ls in 576 288 192; do run variants/b576.so U$ls --recorder uniform --lockstep on --local-size $ls --workload catch_at_age --n 10000000; done
run variants/b672.so TS224 --recorder uniform --lockstep off --local-size 224 --workload tape_stress --n 10000000 --tape-steps 333
run variants/b576.so TS288 --recorder uniform --lockstep off --local-size 288 --workload tape_stress --n 10000000 --tape-steps 333
run variants/b672.so L224 --recorder lane --prefetch on --lockstep on --local-size 224 --workload catch_at_age --n 10000000
