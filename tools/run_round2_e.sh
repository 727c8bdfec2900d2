mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2_kernel -s 244 -c 18 -o gpurun_out/r2_gemm_bwd_mixed python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2_kernel -s 100 -c 10 -o gpurun_out/r2_gemm_fwd_mixed python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
