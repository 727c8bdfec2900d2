# last check of the final code: compute-sanitizer memcheck of the fused SDF chain (ragged size), its test, one C2 line
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 python tools/fused_check.py /tmp/m.pt 70001 > gpurun_out/r2i_memcheck_sdf_fused.txt 2>&1; echo "memcheck fused exit $?"; tail -4 gpurun_out/r2i_memcheck_sdf_fused.txt
timeout 600 python -m pytest tests/test_gpu_sdf_fused.py tests/test_gpu_engine_state.py tests/test_gpu_sdf_grid.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes > gpurun_out/r2i_bench_c2_last.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2i_bench_c2_last.json').read().strip().splitlines()[-1]);print('c2',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['frac'],4),d['clocks'])"
