mkdir -p gpurun_out /tmp/rep; export AD4CL_CUDA_LIB=${LIB:-$PWD/ad4cl_b200/libad4cl_cuda.so}
LIBS=${LIBS:-default}
for rec in ${RECS:-lane uniform}; do
  timeout 600 ncu --section SourceCounters --section InstructionStats --section WarpStateStats --import-source on --clock-control none -k regex:${WL:-catch_at_age} -c 1 -f -o /tmp/rep/src_$rec python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --recorder $rec --prefetch off --workload ${WL:-catch_at_age} --n 10000000 ${WLARGS} > gpurun_out/r02g_ncu_$rec.log 2>&1
  tail -2 gpurun_out/r02g_ncu_$rec.log
  ncu -i /tmp/rep/src_$rec.ncu-rep --page source --csv --print-source sass > gpurun_out/r02g_src_${WL:-catch_at_age}_$rec.csv 2>gpurun_out/r02g_export_$rec.log
  ncu -i /tmp/rep/src_$rec.ncu-rep --page raw --csv > gpurun_out/r02g_raw_${WL:-catch_at_age}_$rec.csv 2>>gpurun_out/r02g_export_$rec.log
done
ls -la gpurun_out/
