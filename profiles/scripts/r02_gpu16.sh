mkdir -p gpurun_out
( time python bench.py --write-golden ) > gpurun_out/r02n_bench_golden.json 2> gpurun_out/r02n_bench_golden.err
tail -5 gpurun_out/r02n_bench_golden.err
( time python bench.py ) > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err
tail -5 gpurun_out/r02n_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02n_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('metric','value','unit','ms_per_step','gpu_launches','parity_check') if k in d})
print('e2e', d['e2e']['value'], 'roofline', {k: d['roofline'][k] for k in ('achieved','peak','frac','traffic','kernel_ms')})
print('cpu_baseline', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('sample'))
for r in d.get('extra', {}).get('configs', []):
    print({k: r.get(k) for k in ('config','workload','kernel_ms','frac','parity_rel_err','value')})
print(d.get('extra', {}).get('config5_sharded'))
PY
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err
tail -4 gpurun_out/r02n_bench_ref.err; cut -c1-600 gpurun_out/r02n_bench_ref.json
cp tests/golden/bench_*.json gpurun_out/ 2>/dev/null; ls tests/golden | head -30
