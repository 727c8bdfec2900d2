mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_optim.py tests/test_gpu_dataio.py -q -m gpu -p no:cacheprovider 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
for i in 1 2; do
NRW_LIB_PATH=$PWD/neuralrecon-w_b200/nrw/libnrw_nol2.so $B > gpurun_out/r2_bench10_nol2_$i.json 2>/dev/null
$B > gpurun_out/r2_bench10_l2_$i.json 2>/dev/null
done
for f in nol2_1 l2_1 nol2_2 l2_2; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench10_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_mixed2.csv python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_step_mixed2.csv 24 > gpurun_out/r2_launches_step_mixed2_summary.txt; cat gpurun_out/r2_launches_step_mixed2_summary.txt; rm -f gpurun_out/r2_launches_step_mixed2.csv
