set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -x 2>&1 | tail -40 > gpurun_out/r02a_tests.log
tail -15 gpurun_out/r02a_tests.log
REPS=1 timeout 600 profiles/ab.sh ad4cl_b200/libad4cl_cuda.so 2>&1 | tee gpurun_out/r02a_ab.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:catch_at_age -c 2 --csv --log-file gpurun_out/r02a_ncu_caa.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_ncu.log 2>&1
tail -5 gpurun_out/r02a_ncu_caa.csv
