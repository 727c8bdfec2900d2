# full-size parity against the unmodified reference on the GPU + compute-sanitizer memcheck of the small paths
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize_reference.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v Warning | tail -6
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck_smoke.txt 2>&1; echo "memcheck smoke exit $?"; tail -5 gpurun_out/r2_memcheck_smoke.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_dataio.py tests/test_gpu_bitexact.py tests/test_gpu_octree_build.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_memcheck_tests.txt 2>&1; echo "memcheck tests exit $?"; tail -6 gpurun_out/r2_memcheck_tests.txt
