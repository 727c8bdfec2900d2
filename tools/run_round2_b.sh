mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider -x --deselect tests/test_gpu_precision_policy.py 2>&1 | tail -30 > gpurun_out/r2_tests2.log
python -m pytest tests/test_gpu_precision_policy.py -q -m gpu -p no:cacheprovider -s 2>&1 | tail -40 > gpurun_out/r2_precision.log
python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/r2_bench2_c2.json 2> gpurun_out/r2_bench2_c2.err
python bench.py --workload C3 --steps 10 --warmup 3 --no_other_modes > gpurun_out/r2_bench2_c3.json 2> gpurun_out/r2_bench2_c3.err
python bench.py --workload C5 --steps 5 --warmup 1 > gpurun_out/r2_bench2_c5.json 2> gpurun_out/r2_bench2_c5.err
tail -8 gpurun_out/r2_tests2.log; tail -25 gpurun_out/r2_precision.log
for f in c2 c3 c5; do echo == $f; cut -c1-300 gpurun_out/r2_bench2_$f.json; tail -3 gpurun_out/r2_bench2_$f.err; done
