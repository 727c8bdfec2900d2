# 4-GPU weak-scaling point of the final code (one short run)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2i_bench_c2_n4.json 2> gpurun_out/r2i_bench_c2_n4.err
python -c "
import json;d=json.loads(open('gpurun_out/r2i_bench_c2_n4.json').read().strip().splitlines()[-1]);print('n4',round(d['ms_per_step'],2),round(d['value']),d['n_gpus'],d['clocks'],d.get('stages'))" || tail -5 gpurun_out/r2i_bench_c2_n4.err
