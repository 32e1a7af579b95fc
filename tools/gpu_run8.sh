mkdir -p gpurun_out
G='"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*'
for m in 1 2 3; do LWDETR_B200_SLOTS_MODE=$m timeout 200 python tools/bench_kernels.py --only window_attention,global_attention --configs small 2>&1 | grep -o "$G" | sed "s/^/SLOTS MODE$m /"; done
for pl in 0 1 2; do LWDETR_B200_ATTN_SLOTS=0 LWDETR_B200_ATTN_TC=0 LWDETR_B200_ATTN_POLY=$pl timeout 200 python tools/bench_kernels.py --only window_attention,global_attention --configs small,medium 2>&1 | grep -o "$G" | sed "s/^/MMASYNC POLY$pl /"; done
LWDETR_B200_SLOTS_MODE=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3
LWDETR_B200_ATTN_SLOTS=0 LWDETR_B200_ATTN_TC=0 LWDETR_B200_ATTN_POLY=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 3 -c 1 -o gpurun_out/r02h_ncu_glb_small python tools/bench_kernels.py --only global_attention --configs small --iters 1 > gpurun_out/r02h_ncu.log 2>&1; tail -1 gpurun_out/r02h_ncu.log
