mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r02h_tests.log 2>&1
tail -12 gpurun_out/r02h_tests.log
for lib in ${LIBS:-ad4cl_b200/libad4cl_cuda.so variants/hash.so variants/outline.so variants/hash_outline.so variants/lane0.so}; do
 for wl in "catch_at_age --n 10000000" "tape_stress --n 10000000 --tape-steps 333"; do
  for rec in "uniform" "lane --prefetch off" "lane --prefetch on"; do
   AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --recorder $rec --workload $wl 2>/dev/null | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$lib', '$wl'.split()[0], '$rec', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], d['launch']['recorder'], d['launch']['sweep_prefetch'])
except Exception as e:
    print('$lib', '$wl'.split()[0], '$rec', 'FAILED', repr(e))
"
  done
 done
done 2>&1 | tee gpurun_out/r02h_ab.log
