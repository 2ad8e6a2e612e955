# Last call of round 2 (4.8 GPU-minutes): the build with deferred re-recording + alternating pass order.  Its ncu
# capture / traffic stamp came back with the previous call (r02_gpu18.sh: 3.191 ms against 3.205 ms at 10^7, gate
# set at 3.19 stopped that call).  Here: the 8-way shard on one GPU with both libraries (where the change is
# expected to matter), the whole GPU suite, and a bench line without the extra configs.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
AD4CL_CUDA_LIB=$PWD/variants/base.so python profiles/scripts/shard_probe.py 8 uniform,on,768 2>&1 | tail -1 | cut -c1-160
python profiles/scripts/shard_probe.py 8 uniform,on,768 2>&1 | tail -1 | cut -c1-160
echo "shard probes done at $(( $(date +%s) - T0 )) s"
timeout 205 python -m pytest tests -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider > gpurun_out/r02b_tests.log 2>&1
tail -12 gpurun_out/r02b_tests.log
echo "tests done at $(( $(date +%s) - T0 )) s"
python bench.py --no-extra --no-cpu-baseline > gpurun_out/r02b_bench_n1_noextra.json 2> gpurun_out/r02b_bench_n1_noextra.err
cut -c1-400 gpurun_out/r02b_bench_n1_noextra.json
echo "all done at $(( $(date +%s) - T0 )) s"
