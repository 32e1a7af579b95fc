mkdir -p gpurun_out
G='"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*'
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -2
LWDETR_B200_SLOTS_MODE=0 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -2
timeout 300 python tools/bench_kernels.py --only window_attention,global_attention 2>&1 | grep -o "$G"
