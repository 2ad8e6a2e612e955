set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r02f_tests.log 2>&1
tail -15 gpurun_out/r02f_tests.log
set +x
for wl in "catch_at_age --n 10000000" "tape_stress --n 10000000 --tape-steps 333"; do
 for rec in "uniform" "lane --prefetch off" "lane --prefetch on" "auto"; do
   timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --recorder $rec --workload $wl 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl'.split()[0], '$rec', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], d['launch']['recorder'], d['launch']['sweep_prefetch'])
"
 done
done 2>&1 | tee gpurun_out/r02f_ab.log
for wl in "catch_at_age --n 10000000" "tape_stress --n 10000000 --tape-steps 333"; do
 for rec in "lane --prefetch off" "lane --prefetch on"; do
   AD4CL_CUDA_LIB=$PWD/variants/lane0.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --recorder $rec --workload $wl 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('LANE0', '$wl'.split()[0], '$rec', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], d['launch']['recorder'], d['launch']['sweep_prefetch'])
"
 done
done 2>&1 | tee -a gpurun_out/r02f_ab.log
REPS=1 RECS=auto bash profiles/ab.sh ad4cl_b200/libad4cl_cuda.so 2>&1 | tee -a gpurun_out/r02f_ab.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
for s in long_scoreboard short_scoreboard wait no_instruction math_pipe_throttle branch_resolving barrier not_selected; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
for rec in uniform lane; do
 for wl in "catch_at_age --n 10000000" "tape_stress --n 10000000 --tape-steps 333"; do
  n=$(echo $wl | cut -d' ' -f1)
  timeout 300 ncu --metrics $M --clock-control none -k regex:$n -c 1 --csv --log-file gpurun_out/r02f_ncu_${n}_$rec.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --recorder $rec --prefetch off --workload $wl > gpurun_out/r02f_ncu.log 2>&1
  echo == $n $rec; grep -v "^==" gpurun_out/r02f_ncu_${n}_$rec.csv | awk -F'","' '{print $13, $15}' | grep -v "Metric Name"
 done
done
