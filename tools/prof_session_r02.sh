# Round-2 profiling session (run on the GPU box through gpurun; summaries are made in the build container with
# tools/ncu_summary.py and committed under profiles/). Numbers printed by a run under ncu are never bench values.
set -x
KRE='regex:tcgen05|norm_kernel|im2col|pool_norm|build_lm|resample|score_|topk_rows|exact_scores|f32_to_f16|rescore'
# (1) launch list of two timed device-resident steps of the bench (kernel shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 1425 -c 950 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-torch-baseline --small-batch 0 > gpurun_out/r02_launches_bench.log 2>&1
tail -c 300 gpurun_out/r02_launches_bench.log
# (2) ncu --set full of every non-GEMM kernel at the bench shapes (second launch of each = warm)
ncu --set full --clock-control none --import-source on -k "regex:norm_kernel|im2col|pool_norm|build_lm|score_filter|rescore|topk_rows|exact_scores|f32_to_f16" \
    -f -o gpurun_out/r02_prof_kernels python tools/bench_kernels.py --ncu > gpurun_out/r02_prof_kernels.log 2>&1
tail -3 gpurun_out/r02_prof_kernels.log
ls -la gpurun_out/*.ncu-rep
