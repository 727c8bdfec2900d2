python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 8 > gpurun_out/bench_r1_full.json 2> gpurun_out/bench_r1_full.err; tail -c 3000 gpurun_out/bench_r1_full.json
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-400
# top-kernel capture inside the real training step: skip the first 150 gemm_tc2 launches, capture 2
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2_kernel -s 150 -c 2 -o gpurun_out/gemm2_r1_step python tools/prof_step.py --chunk_rows 262144 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
