N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_sharded_local.py -m gpu -q --tb=short > gpurun_out/r02_multi_tests_n$N.log 2>&1
tail -15 gpurun_out/r02_multi_tests_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -5 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('metric','value','unit','n_gpus','ms_per_step','gpu_launches','parity_check','scaling')})
print('e2e', d['e2e'], 'kernel_ms', d['roofline']['kernel_ms'], d['roofline'].get('kernel_ms_by_rank'))
print('launch', d['launch'])
print('config5', json.dumps(d.get('extra', {}).get('config5_sharded'))[:1500])
PY
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus $N --steps 1 --warmup 1 ) > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err
cut -c1-300 gpurun_out/r02_bench_ref_n$N.json
