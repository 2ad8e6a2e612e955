# 1.2 GPU-minutes: rank 0's shard of the 8-way problem on one GPU, four libraries, alternating, twice.
for round in 1 2; do
for lib in base snake defer both; do
  if [ $lib = both ]; then L=$PWD/ad4cl_b200/libad4cl_cuda.so; else L=$PWD/variants/$lib.so; fi
  AD4CL_CUDA_LIB=$L PROBE_EVALS=60 timeout 15 python profiles/scripts/shard_probe.py 8 uniform,on,768 2>&1 | tail -1 | cut -c1-70 | sed "s/^/$lib /"
done; done
