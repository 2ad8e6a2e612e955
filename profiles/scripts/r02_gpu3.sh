set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short > gpurun_out/r02d_tests.log 2>&1
tail -25 gpurun_out/r02d_tests.log
REPS=1 timeout 900 profiles/ab.sh ad4cl_b200/libad4cl_cuda.so variants/hash.so ad4cl_b200/libad4cl_cuda_general.so 2>&1 | tee gpurun_out/r02d_ab.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
for s in long_scoreboard short_scoreboard wait no_instruction math_pipe_throttle mio_throttle lg_throttle branch_resolving barrier not_selected dispatch_stall; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
set +x
timeout 300 ncu --metrics $M --clock-control none -k regex:catch_at_age -c 1 --csv --log-file gpurun_out/r02d_ncu_caa.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02d_ncu.log 2>&1
grep -v "^==" gpurun_out/r02d_ncu_caa.csv | awk -F'","' '{print $13, $15}'
AD4CL_CUDA_LIB=$PWD/variants/hash.so timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:catch_at_age -c 1 --csv --log-file gpurun_out/r02d_ncu_caa_hash.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02d_ncu2.log 2>&1
grep -v "^==" gpurun_out/r02d_ncu_caa_hash.csv | awk -F'","' '{print $13, $15}'
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tape_stress -c 1 --csv --log-file gpurun_out/r02d_ncu_ts.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --workload tape_stress --n 10000000 --tape-steps 333 > gpurun_out/r02d_ncu3.log 2>&1
grep -v "^==" gpurun_out/r02d_ncu_ts.csv | awk -F'","' '{print $13, $15}'
