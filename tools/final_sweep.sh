# final evidence of the round: GPU tests, smoke, both bench arms (no ncu)
python -m pytest tests -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-700 gpurun_out/bench_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_final_reference.json
