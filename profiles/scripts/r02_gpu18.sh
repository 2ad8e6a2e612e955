# One call, 5.7 GPU-minutes left: the build with deferred re-recording + alternating pass order (ad4cl_entry.cuh)
# against the previous build (variants/base.so).  Gate on the headline kernel time; then the 8-way shard on one
# GPU with both libraries, the GPU suite, and a final bench line without the extra configs.
set -u
mkdir -p gpurun_out /tmp/rep gpurun_out/profiles
TAG=r02b
T0=$(date +%s)
cap() { # name, kernel-regex, traffic-key, bench args...
  name=$1; rx=$2; key=$3; shift 3
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" > gpurun_out/${TAG}_${name}_plain.json 2> gpurun_out/${TAG}_${name}_plain.err || { echo "plain run of $name failed"; tail -5 gpurun_out/${TAG}_${name}_plain.err; return 1; }
  FORCE=$(python -c "
import json, sys
l = json.loads(open('gpurun_out/${TAG}_${name}_plain.json').read().strip().splitlines()[-1])['launch']
if l['tape_tier'] != 'register':
    print('--recorder', l['recorder'], '--prefetch', 'on' if l['sweep_prefetch'] == 1 else 'off', '--lockstep', 'on' if l['lockstep'] else 'off', '--local-size', l['block'])
")
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:$rx -c 1 -f -o /tmp/rep/$name python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra "$@" $FORCE > gpurun_out/${TAG}_${name}_ncu.log 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2> gpurun_out/${TAG}_${name}_export.log
  python profiles/tools/ncu_summary.py gpurun_out/${TAG}_${name}_raw.csv gpurun_out/profiles/${TAG}_ncu_${name}.json --traffic-key "$key" --traffic-file gpurun_out/profiles/dram_traffic.json --source "profiles/${TAG}_ncu_${name}.json (ncu --set full, one launch: bench.py $* $FORCE)"
}
cp profiles/dram_traffic.json gpurun_out/profiles/dram_traffic.json
cap catch_at_age_10m catch_at_age "catch_at_age:10000000:1" --workload catch_at_age --n 10000000 || exit 1
python - <<'PY' || { echo "GATE: no gain at 10^7, stopping to keep the GPU minutes"; exit 0; }
import json, sys
d = json.loads(open('gpurun_out/r02b_catch_at_age_10m_plain.json').read().strip().splitlines()[-1])
ms = d['roofline']['kernel_ms']
print('new build: kernel_ms', ms, 'value', d['value'], 'parity', d['parity_check']['status'], d['parity_check']['rel_err_f'], d['parity_check']['rel_err_grad'], d['launch']['recorder'], d['launch']['block'], d['launch']['lockstep'])
sys.exit(0 if ms < 3.19 and d['parity_check']['status'] == 'ok' else 1)
PY
echo "cap + gate done at $(( $(date +%s) - T0 )) s"
AD4CL_CUDA_LIB=$PWD/variants/base.so python profiles/scripts/shard_probe.py 8 uniform,on,768 2>&1 | tail -1 | cut -c1-200
python profiles/scripts/shard_probe.py 8 uniform,on,768 2>&1 | tail -1 | cut -c1-200
echo "shard probes done at $(( $(date +%s) - T0 )) s"
timeout 215 python -m pytest tests -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider > gpurun_out/r02b_tests.log 2>&1
tail -12 gpurun_out/r02b_tests.log
echo "tests done at $(( $(date +%s) - T0 )) s"
cp gpurun_out/profiles/dram_traffic.json profiles/dram_traffic.json
python bench.py --no-extra --no-cpu-baseline > gpurun_out/r02b_bench_n1_noextra.json 2> gpurun_out/r02b_bench_n1_noextra.err
cut -c1-600 gpurun_out/r02b_bench_n1_noextra.json
echo "all done at $(( $(date +%s) - T0 )) s"
