# Final evidence run of round 2 for the build in the tree: GPU suite, smoke, both bench arms, ncu summaries.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r02_final_tests.log 2>&1
tail -6 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','parity_check')}, 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'traffic', d['roofline']['traffic'], d['launch']['recorder'], d['launch']['block'], d['launch']['lockstep'])
for r in d['extra']['configs']: print(r['workload'], round(r['kernel_ms'], 4), r['parity_rel_err'])
for r in d['extra']['config5_at_1e8']: print(r['workload'], round(r['ms_per_eval'], 2), r['parity_rel_err'])
r = json.loads(open('gpurun_out/r02_bench_reference_n1.json').read().strip().splitlines()[-1]); print('reference', r['value'], r.get('scaled_from_sample'))
PY
bash profiles/scripts/r02_profile.sh
python examples/bench_dense.py 1024 2048 4096 > gpurun_out/r02_dense.json; cut -c1-400 gpurun_out/r02_dense.json
