# fused forward-only SDF chain: correctness vs the per-layer chain (small, ragged, multi-chunk), then timing
mkdir -p gpurun_out
for n in 1000 300001 2097152; do
NRW_SDF_FUSED=0 timeout -s KILL 120 python tools/fused_check.py /tmp/u_$n.pt $n 2>&1 | tail -1
NRW_SDF_FUSED=1 timeout -s KILL 120 python tools/fused_check.py /tmp/f_$n.pt $n 2>&1 | tail -2
timeout 60 python tools/fused_check.py cmp /tmp/u_$n.pt /tmp/f_$n.pt 2>&1 | tail -4
done
NRW_SDF_FUSED=0 timeout -s KILL 300 python bench.py --workload C5 --dim 256 --steps 5 --warmup 2 --no_torch_gpu_ref > gpurun_out/r2_fz_c5_256_off.json 2>/dev/null
NRW_SDF_FUSED=1 timeout -s KILL 300 python bench.py --workload C5 --dim 256 --steps 5 --warmup 2 --no_torch_gpu_ref > gpurun_out/r2_fz_c5_256_on.json 2>/dev/null
for f in off on; do python -c "
import json;d=json.loads(open('gpurun_out/r2_fz_c5_256_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),d['clocks'], d.get('sdf_min_max'))"; done
