mkdir -p gpurun_out
./tools/ubench/softmax_rate > gpurun_out/r02g_ubench_softmax.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 2 -c 2 -o gpurun_out/r02g_ncu_attn_small python tools/bench_kernels.py --only window_attention,global_attention --configs small --iters 1 > gpurun_out/r02g_ncu_attn.log 2>&1; tail -2 gpurun_out/r02g_ncu_attn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 1 -o gpurun_out/r02g_ncu_msda_medium python tools/bench_kernels.py --only msda_forward --configs medium --iters 3 > gpurun_out/r02g_ncu_msda.log 2>&1; tail -2 gpurun_out/r02g_ncu_msda.log
for s in 0; do LWDETR_B200_ATTN_SLOTS=$s timeout 300 python tools/bench_kernels.py --only window_attention,global_attention --configs small,medium 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*' | sed "s/^/SLOTS$s /"; done
cat gpurun_out/r02g_ubench_softmax.txt
