mkdir -p gpurun_out
run() { # label, bench args...
  label=$1; shift 1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/err.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'ms_per_step=%.4f' % d['ms_per_step'], 'evals/s=%.2f' % d['value'], l['grid'], l['block'], l['recorder'], 'pf', l['sweep_prefetch'], 'ls', l['lockstep'])
except Exception as e:
    print('$label', 'FAILED', repr(e)); print(open('gpurun_out/err.log').read()[-600:])
"
}
for n in 113664 1250000 10000000; do
for fl in "" "--no-l2-flush"; do
run U768ls-$n$fl --recorder uniform --lockstep on --local-size 768 --workload catch_at_age --n $n $fl
run L256-pf1-$n$fl --recorder lane --prefetch on --lockstep off --local-size 256 --workload catch_at_age --n $n $fl
done
done 2>&1 | tee gpurun_out/r02_small2.log
