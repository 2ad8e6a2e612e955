# Last GPU call of round 2 (11 GPU-minutes left): the device headers changed after the previous capture only inside
# "#if AD4CL_DEBUG_BOUNDS" blocks, but the device_abi hash covers the header text, so the DRAM-traffic stamp of the
# headline launch is re-captured for the build in the tree; then the default bench line (now with `traffic`), smoke,
# and as much of the GPU suite as the remaining budget allows.
set -u
mkdir -p gpurun_out /tmp/rep gpurun_out/profiles
TAG=${TAG:-r02}
cap() { # name, kernel-regex, traffic-key, bench args...
  name=$1; rx=$2; key=$3; shift 3
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra "$@" > gpurun_out/${TAG}_${name}_plain.json 2> gpurun_out/${TAG}_${name}_plain.err || { echo "plain run of $name failed"; return; }
  FORCE=$(python -c "
import json, sys
l = json.loads(open('gpurun_out/${TAG}_${name}_plain.json').read().strip().splitlines()[-1])['launch']
if l['tape_tier'] != 'register':
    print('--recorder', l['recorder'], '--prefetch', 'on' if l['sweep_prefetch'] == 1 else 'off', '--lockstep', 'on' if l['lockstep'] else 'off', '--local-size', l['block'])
")
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$rx -c 1 -f -o /tmp/rep/$name python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra "$@" $FORCE > gpurun_out/${TAG}_${name}_ncu.log 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2> gpurun_out/${TAG}_${name}_export.log
  python profiles/tools/ncu_summary.py gpurun_out/${TAG}_${name}_raw.csv gpurun_out/profiles/${TAG}_ncu_${name}.json --traffic-key "$key" --traffic-file gpurun_out/profiles/dram_traffic.json --source "profiles/${TAG}_ncu_${name}.json (ncu --set full, one launch: bench.py $* $FORCE)"
}
T0=$(date +%s)
cp profiles/dram_traffic.json gpurun_out/profiles/dram_traffic.json
cap catch_at_age_10m catch_at_age "catch_at_age:10000000:1" --workload catch_at_age --n 10000000
cp gpurun_out/profiles/dram_traffic.json profiles/dram_traffic.json
echo "cap done at $(( $(date +%s) - T0 )) s"
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','gpu_launches','parity_check')}, 'e2e', d['e2e']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'traffic', d['roofline']['traffic'], d['roofline']['dram_frac'], d['launch']['recorder'], d['launch']['block'], d['launch']['lockstep'])
for r in d['extra']['configs']: print(r['workload'], round(r['kernel_ms'], 4), r['parity_rel_err'])
PY
echo "bench done at $(( $(date +%s) - T0 )) s"
python -c "import __graft_entry__ as g; g.smoke()"
timeout ${PYTEST_BUDGET:-300} python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/r02_final_tests.log 2>&1
tail -8 gpurun_out/r02_final_tests.log
echo "tests done at $(( $(date +%s) - T0 )) s"
cap regression_linear_1m regression_linear "regression_linear:1000000:1" --workload regression_linear --n 1000000
cap linreg2_1m linreg2 "linreg2_simple:1000000:1" --workload linreg2_simple --n 1000000
echo "all done at $(( $(date +%s) - T0 )) s"
ls -la gpurun_out/profiles
