mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_precision_policy.py 2>&1 | tail -15 > gpurun_out/r2_tests4.log
tail -4 gpurun_out/r2_tests4.log
rm -f /tmp/g.csv
NRW_GEMM_TIMING_DUMP=/tmp/g.csv python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes > gpurun_out/r2_bench4_mixed.json 2> gpurun_out/r2_bench4_mixed.err
python tools/gemm_table.py /tmp/g.csv > gpurun_out/r2_gemm_table_mixed.txt 2>&1; head -30 gpurun_out/r2_gemm_table_mixed.txt
NRW_BWD_GATE_PLANES=1 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes > gpurun_out/r2_bench4_mixed_gate1.json 2> gpurun_out/r2_bench4_mixed_gate1.err
python bench.py --precision bf16x3 --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes > gpurun_out/r2_bench4_bf16x3.json 2> gpurun_out/r2_bench4_bf16x3.err
for f in mixed mixed_gate1 bf16x3; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench4_$f.json').read().strip().splitlines()[-1]);print('$f',d['ms_per_step'],d['value'],d['forward_slots'],d['roofline']['frac'],d['roofline']['kernel_ms_per_step'],d['roofline']['algorithmic_hbm_gb_per_step_in_kernel'],d['clocks']['sm_mhz'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_mixed.csv python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_step_mixed.csv 40 > gpurun_out/r2_launches_step_mixed_summary.txt; head -30 gpurun_out/r2_launches_step_mixed_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"head_bwd|sdf_normal_bwd|nerf_embed|sdf_head|composite|upsample_round|sdf_embed|color_embed|colsum|head_kernel|sdf_normal_kernel" -c 16 -o gpurun_out/r2_pointwise python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
