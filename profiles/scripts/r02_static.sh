mkdir -p gpurun_out
run() { # lib label, bench args...
  lib=$1; label=$2; shift 2
  AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/err.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$lib', '$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'ms_per_step=%.4f' % d['ms_per_step'], 'evals/s=%.2f' % d['value'], 'e2e=%.1f' % d['e2e']['value'], l['grid'], l['block'], l['tape_tier'])
except Exception as e:
    print('$lib', '$label', 'FAILED', repr(e)); print(open('gpurun_out/err.log').read()[-400:])
"
}
L=variants/fin.so
for wl in linreg2_simple regression_linear regression_logistic; do
for bps in 0 1 2 3 4; do
run $L $wl-bps$bps --workload $wl --n 1000000 --blocks-per-sm $bps
done
for ls in 128 64; do
run $L $wl-ls$ls --workload $wl --n 1000000 --local-size $ls
done
done
