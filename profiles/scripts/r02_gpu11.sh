mkdir -p gpurun_out
run() { # lib, label, bench args...
  lib=$1; label=$2; shift 2
  AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" 2>/dev/null | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$lib', '$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], l['grid'], l['block'], l['smem_bytes'], l['recorder'], l['sweep_prefetch'])
except Exception as e:
    print('$lib', '$label', 'FAILED', repr(e))
"
}
for lib in variants/lockstep.so variants/lockstep2.so; do
 for ls in 256 512 768 128; do
  run $lib caa-uniform-$ls --recorder uniform --local-size $ls --workload catch_at_age --n 10000000
  run $lib caa-lane-pf1-$ls --recorder lane --prefetch on --local-size $ls --workload catch_at_age --n 10000000
 done
 run $lib ts-uniform --recorder uniform --workload tape_stress --n 10000000 --tape-steps 333
 run $lib ts-uniform-768 --recorder uniform --local-size 768 --workload tape_stress --n 10000000 --tape-steps 333
done 2>&1 | tee gpurun_out/r02j_ab.log
