# re-check after the compile-time schedule refactor of the fused chain (scalar index code only)
timeout 600 python -m pytest tests/test_gpu_sdf_fused.py tests/test_gpu_bitexact.py tests/test_gpu_sdf_grid.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout -s KILL 120 python tools/fused_check.py /tmp/f.pt 2097152 2>&1 | tail -1
