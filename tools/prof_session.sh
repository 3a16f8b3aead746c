set -x
KRE='regex:tcgen05|norm_kernel|im2col|pool_norm|build_lm|resample|score_|topk_rows|exact_scores|f32_to_f16'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 1425 -c 950 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -c 400 gpurun_out/launches_bench.log
PROF_SLICES=128 ncu --set full --clock-control none --import-source on -k "regex:gemm2?_tcgen05" -s 4 -c 4 -f -o gpurun_out/prof_gemm_r01c python tools/prof_kernels.py gemm > gpurun_out/ncu_gemm_c.log 2>&1
PROF_SLICES=128 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 2 -c 1 -f -o gpurun_out/prof_attn_r01c python tools/prof_kernels.py attn > gpurun_out/ncu_attn_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
