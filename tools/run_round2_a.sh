mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2_tests1.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_tests1.log; cat gpurun_out/r2_ref1.json | cut -c1-400; cat gpurun_out/r2_bench1.json | cut -c1-600
