mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_c2_n2.json 2> gpurun_out/r2f_bench_c2_n2.err
tail -3 gpurun_out/r2f_bench_c2_n2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_c2_n2.json').read().strip().splitlines()[-1])
print('n2', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', d['n_gpus'], 'e2e', round(d['e2e']['value'],1), d['config']['parallelism'], d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2f_bench_c5_n2.json 2> gpurun_out/r2f_bench_c5_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_c5_n2.json').read().strip().splitlines()[-1])
print('c5 n2', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', d['n_gpus'])
PY
