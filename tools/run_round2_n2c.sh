mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_c2_n2.json 2> gpurun_out/r2i_bench_c2_n2.err
tail -3 gpurun_out/r2i_bench_c2_n2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_bench_c2_n2.json').read().strip().splitlines()[-1])
print('n2', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', d['n_gpus'], 'e2e', round(d['e2e']['value'],1), d['config']['parallelism'], d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2i_bench_c5_n2.json 2> gpurun_out/r2i_bench_c5_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_bench_c5_n2.json').read().strip().splitlines()[-1])
print('c5 n2', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', d['n_gpus'])
PY
# single-GPU reference point on the SAME box (GPU 0), so the 2-GPU efficiency is not a box-to-box comparison
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes > gpurun_out/r2i_bench_c2_n1_samebox.json 2>/dev/null
python - <<PY
import json
for f in ('c2_n2','c2_n1_samebox'):
    d=json.loads(open('gpurun_out/r2i_bench_%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],2), round(d['value']), d['clocks'], d.get('stages'))
PY
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes > gpurun_out/r2i_bench_c2_n1_samebox_gpu1.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/r2i_bench_c2_n1_samebox_gpu1.json').read().strip().splitlines()[-1]); print('gpu1 alone', round(d['ms_per_step'],2), round(d['value']), d['clocks'])"
