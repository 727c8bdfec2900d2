mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_precision_policy.py 2>&1 | tail -60 > gpurun_out/r2_tests3.log
python -m pytest tests/test_gpu_precision_policy.py -q -m gpu -p no:cacheprovider -s 2>&1 | tail -80 > gpurun_out/r2_precision.log
python bench.py --workload C3 --steps 10 --warmup 3 --no_other_modes > gpurun_out/r2_bench3_c3.json 2> gpurun_out/r2_bench3_c3.err
tail -30 gpurun_out/r2_tests3.log; tail -60 gpurun_out/r2_precision.log
cut -c1-300 gpurun_out/r2_bench3_c3.json; tail -3 gpurun_out/r2_bench3_c3.err
