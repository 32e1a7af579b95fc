mkdir -p gpurun_out
G='"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*'
LWDETR_B200_SLOTS_MODE=3 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3
LWDETR_B200_SLOTS_MODE=0 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3
for m in 0 1 2 3; do LWDETR_B200_SLOTS_MODE=$m timeout 200 python tools/bench_kernels.py --only window_attention,global_attention --configs small,medium 2>&1 | grep -o "$G" | sed "s/^/SLOTS MODE$m /"; done
LWDETR_B200_ATTN_SLOTS=2 LWDETR_B200_SLOTS_MODE=1 timeout 200 python tools/bench_kernels.py --only global_attention --configs medium,large 2>&1 | grep -o "$G" | sed "s/^/SLOTS2 MODE1 /"
LWDETR_B200_ATTN_SLOTS=2 LWDETR_B200_SLOTS_MODE=0 timeout 200 python tools/bench_kernels.py --only global_attention --configs medium,large 2>&1 | grep -o "$G" | sed "s/^/SLOTS2 MODE0 /"
LWDETR_B200_SLOTS_MODE=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 3 -c 1 -o gpurun_out/r02i_ncu_glb_small python tools/bench_kernels.py --only global_attention --configs small --iters 1 > gpurun_out/r02i_ncu.log 2>&1; tail -1 gpurun_out/r02i_ncu.log
