mkdir -p gpurun_out
run() { # lib label, bench args...
  lib=$1; label=$2; shift 2
  AD4CL_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/err.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    l = d['launch']
    print('$lib', '$label', 'kernel_ms=%.4f' % d['roofline']['kernel_ms'], 'evals/s=%.2f' % d['value'], 'e2e=%.2f' % d['e2e']['value'], l['grid'], l['block'], l['smem_bytes'], l['tape_tier'], l['recorder'], 'pf', l['sweep_prefetch'], 'ls', l['lockstep'])
except Exception as e:
    print('$label', 'FAILED', repr(e)); print(open('gpurun_out/err.log').read()[-600:])
"
}
for lib in ad4cl_b200/libad4cl_cuda.so variants/outline.so; do
run $lib caa-auto --workload catch_at_age --n 10000000
run $lib caa-U256 --recorder uniform --lockstep on --local-size 256 --workload catch_at_age --n 10000000
run $lib caa-L256 --recorder lane --lockstep on --prefetch on --local-size 256 --workload catch_at_age --n 10000000
done 2>&1 | tee gpurun_out/r02m_ab.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
for s in long_scoreboard short_scoreboard wait no_instruction math_pipe_throttle branch_resolving barrier not_selected mio_throttle lg_throttle dispatch_stall; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
timeout 300 ncu --metrics $M --clock-control none -k regex:catch_at_age -c 1 --csv --log-file gpurun_out/r02m_ncu_auto.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --recorder uniform --lockstep on --workload catch_at_age --n 10000000 > gpurun_out/r02m_ncu.log 2>&1
echo == auto; grep -v "^==" gpurun_out/r02m_ncu_auto.csv | awk -F'","' '{print $13, $15}' | grep -v "Metric Name"
