mkdir -p gpurun_out
G='"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*'
for sp in 1 2; do LWDETR_B200_ATTN_SLOTS=$sp timeout 200 python tools/bench_kernels.py --only window_attention,global_attention 2>&1 | grep -o "$G" | sed "s/^/SLOTS$sp /"; done
LWDETR_B200_ATTN_SLOTS=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 4 -c 2 -o gpurun_out/r02j_ncu_attn_small python tools/bench_kernels.py --only window_attention,global_attention --configs small --iters 2 > gpurun_out/r02j_ncu.log 2>&1; tail -1 gpurun_out/r02j_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 4 -o gpurun_out/r02j_ncu_gemm_small python tools/one_forward.py --config small --batch 32 --n 1 > gpurun_out/r02j_ncu_gemm.log 2>&1; tail -1 gpurun_out/r02j_ncu_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r02j_ops_small.json > gpurun_out/r02j_bench_small.log 2>&1; tail -c 3000 gpurun_out/r02j_bench_small.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 1 -o gpurun_out/r02j_ncu_msda_medium python tools/bench_kernels.py --only msda_forward --configs medium --iters 3 > gpurun_out/r02j_ncu_msda.log 2>&1; tail -1 gpurun_out/r02j_ncu_msda.log
