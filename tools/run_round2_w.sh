# fused SDF chain v3 (per-k-block readiness): run-to-run determinism, the GPU suite with the fused chain as default, ncu --set full
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout -s KILL 120 python tools/fused_check.py /tmp/f1.pt 2097152 2>&1 | tail -1
timeout -s KILL 120 python tools/fused_check.py /tmp/f2.pt 2097152 2>&1 | tail -1
timeout 60 python tools/fused_check.py cmp /tmp/f1.pt /tmp/f2.pt 2>&1 | tail -3
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sdf_fused" -s 1 -c 1 -o gpurun_out/r2h_sdf_fused python tools/fused_check.py /tmp/x.pt 2097152 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
