# where do the 6 % of the 2-GPU step go?  (a) two INDEPENDENT single-GPU processes at the same time, (b) torchrun with the
# all-reduce, (c) torchrun without it (ranks unsynchronised), (d) host issue time of one step
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
CUDA_VISIBLE_DEVICES=0 timeout 300 $B > gpurun_out/y_ind0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 timeout 300 $B > gpurun_out/y_ind1.json 2>/dev/null &
wait
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$T 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/y_ddp.json 2>/dev/null
NRW_DIAG_SKIP_REDUCE=1 $T 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/y_ddp_noreduce.json 2>/dev/null
NRW_PDL=0 $T 29523 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/y_ddp_nopdl.json 2>/dev/null
for f in ind0 ind1 ddp ddp_noreduce ddp_nopdl; do python -c "
import json;d=json.loads(open('gpurun_out/y_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),d['clocks'].get('sm_mhz'),d['clocks'].get('power_w'),d.get('stages',{}).get('compute_ms_per_rank'),d.get('stages',{}).get('reduce_ms_per_rank'))"; done
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/step_timing.py 2>&1 | grep -E "one step|async 20|sync each"
