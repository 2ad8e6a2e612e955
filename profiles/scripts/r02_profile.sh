# One gpurun call: the ncu evidence of round 2 for the build in the tree.  Each capture runs only after the
# same command has exited 0 without ncu.  Reports stay on the box; the raw/source pages come back as CSV in
# gpurun_out/ and the summaries are written straight into profiles/ of the box copy and copied back.
set -u
mkdir -p gpurun_out /tmp/rep gpurun_out/profiles
TAG=${TAG:-r02}
cap() { # name, kernel-regex, traffic-key, bench args...
  name=$1; rx=$2; key=$3; shift 3
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra "$@" > gpurun_out/${TAG}_${name}_plain.json 2> gpurun_out/${TAG}_${name}_plain.err || { echo "plain run of $name failed"; return; }
  # the tuner's choice of the plain run, forced for the capture (under ncu the first matching launch would be the tuner's first candidate)
  FORCE=$(python -c "
import json, sys
l = json.loads(open('gpurun_out/${TAG}_${name}_plain.json').read().strip().splitlines()[-1])['launch']
if l['tape_tier'] != 'register':
    print('--recorder', l['recorder'], '--prefetch', 'on' if l['sweep_prefetch'] == 1 else 'off', '--lockstep', 'on' if l['lockstep'] else 'off', '--local-size', l['block'])
")
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$rx -c 1 -f -o /tmp/rep/$name python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra "$@" $FORCE > gpurun_out/${TAG}_${name}_ncu.log 2>&1
  ncu -i /tmp/rep/$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2> gpurun_out/${TAG}_${name}_export.log
  ncu -i /tmp/rep/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_${name}_source.csv 2>> gpurun_out/${TAG}_${name}_export.log
  python profiles/tools/ncu_summary.py gpurun_out/${TAG}_${name}_raw.csv gpurun_out/profiles/${TAG}_ncu_${name}.json --traffic-key "$key" --traffic-file gpurun_out/profiles/dram_traffic.json --source "profiles/${TAG}_ncu_${name}.json (ncu --set full, one launch: bench.py $* $FORCE)"
}
cp profiles/dram_traffic.json gpurun_out/profiles/dram_traffic.json
cap catch_at_age_10m catch_at_age "catch_at_age:10000000:1" --workload catch_at_age --n 10000000
cap tape_stress_10m_L1000 tape_stress "tape_stress:10000000:1" --workload tape_stress --n 10000000 --tape-steps 333
cap tape_stress_2m_L10000 tape_stress "tape_stress:2000000:1:3333" --workload tape_stress --n 2000000 --tape-steps 3333
cap regression_linear_1m regression_linear "regression_linear:1000000:1" --workload regression_linear --n 1000000
cap linreg2_1m linreg2 "linreg2_simple:1000000:1" --workload linreg2_simple --n 1000000
# launch list of the default bench command (every kernel of a short run, cold-cache serialised times)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_launches_plain.json 2>/dev/null && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/profiles/${TAG}_launches_catch_at_age_10m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_launches_ncu.log 2>&1
ls -la gpurun_out/profiles
