mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r02l_tests.log 2>&1
tail -8 gpurun_out/r02l_tests.log
run() { # label, bench args...
  label=$1; shift 1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/err.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'ms_per_step=%.4f' % d['ms_per_step'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], l['grid'], l['block'], l['smem_bytes'], l['tape_tier'], l['recorder'], 'pf', l['sweep_prefetch'], 'ls', l['lockstep'])
except Exception as e:
    print('$label', 'FAILED', repr(e)); print(open('gpurun_out/err.log').read()[-600:])
"
}
for rep in 1 2; do
run caa-auto --workload catch_at_age --n 10000000
run ts-auto --workload tape_stress --n 10000000 --tape-steps 333
run reglin-auto --workload regression_linear --n 1000000
run reglog-auto --workload regression_logistic --n 1000000
run linreg2-auto --workload linreg2_simple --n 1000000
done 2>&1 | tee gpurun_out/r02l_ab.log
run caa-U-ls-off --recorder uniform --lockstep off --workload catch_at_age --n 10000000 | tee -a gpurun_out/r02l_ab.log
run caa-U-ls-on --recorder uniform --lockstep on --workload catch_at_age --n 10000000 | tee -a gpurun_out/r02l_ab.log
run caa-U-256 --recorder uniform --lockstep on --local-size 256 --workload catch_at_age --n 10000000 | tee -a gpurun_out/r02l_ab.log
run caa-n1e6 --workload catch_at_age --n 1000000 | tee -a gpurun_out/r02l_ab.log
run caa-n1e5 --workload catch_at_age --n 100000 | tee -a gpurun_out/r02l_ab.log
