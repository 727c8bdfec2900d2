mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"gemm_tc2_kernel<0, 6>" -s 8 -c 2 -o gpurun_out/r2_k6_reverse python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"gemm_tc2_kernel<0, 4>" -s 8 -c 2 -o gpurun_out/r2_k4_gatefwd python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"gemm_tc2_kernel<0, 1>" -s 60 -c 1 -o gpurun_out/r2_k1_fwd python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"composite|merge_sorted|points_kernel|coarse_z" -c 6 -o gpurun_out/r2_composite python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
