# round-1 evidence sweep (one B200): tests, bench (both arms), ncu launch list of one step, ncu --set full of the top kernel
python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --steps 8 > gpurun_out/bench_r1_full.json 2> gpurun_out/bench_r1_full.err; tail -c 2500 gpurun_out/bench_r1_full.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r1_reference.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
# launch list of ONE training step (cudaProfilerStart/Stop range in tools/prof_step.py)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step.csv python tools/prof_step.py --chunk_rows 262144 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_step.csv 14
# top kernel inside the real step: skip the first 150 gemm_tc2 launches, capture 2
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2_kernel -s 150 -c 2 -o gpurun_out/gemm2_r1b_step python tools/prof_step.py --chunk_rows 262144 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
