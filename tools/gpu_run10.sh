mkdir -p gpurun_out
G='"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*'
LWDETR_B200_SLOTS_MODE=3 timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -3
LWDETR_B200_SLOTS_MODE=0 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -3
timeout 200 python tools/bench_kernels.py --only msda_forward 2>&1 | grep -o "$G" | sed "s/^/MSDA /"
for m in 0 1 2 3; do LWDETR_B200_SLOTS_MODE=$m timeout 200 python tools/bench_kernels.py --only window_attention,global_attention --configs small 2>&1 | grep -o "$G" | sed "s/^/MODE$m POLY3 /"; done
for pl in 0 1 2; do LWDETR_B200_SLOTS_MODE=3 LWDETR_B200_SLOTS_POLY=$pl timeout 200 python tools/bench_kernels.py --only window_attention,global_attention --configs small,medium 2>&1 | grep -o "$G" | sed "s/^/MODE3 POLY$pl /"; done
for pl in 1 2; do LWDETR_B200_SLOTS_MODE=2 LWDETR_B200_SLOTS_POLY=$pl timeout 200 python tools/bench_kernels.py --only global_attention --configs small 2>&1 | grep -o "$G" | sed "s/^/MODE2 POLY$pl /"; done
