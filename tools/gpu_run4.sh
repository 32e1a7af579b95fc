mkdir -p gpurun_out
timeout 120 python tools/attn_debug.py 2>&1 | tail -20
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r02d_test_attn.log 2>&1; tail -8 gpurun_out/r02d_test_attn.log
timeout 300 python tools/bench_kernels.py --only window_attention,global_attention --out gpurun_out/r02d_kernels_attn.json 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*' 
LWDETR_B200_ATTN_SLOTS=2 timeout 300 python tools/bench_kernels.py --only global_attention --configs medium,large 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*' | sed 's/^/SLOTS2 /'
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 2 -o gpurun_out/r02d_ncu_msda_small python tools/bench_kernels.py --only msda_forward --configs small --iters 3 > gpurun_out/r02d_ncu_msda.log 2>&1; tail -3 gpurun_out/r02d_ncu_msda.log
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "not baseline and not medium and not xlarge and not large" > gpurun_out/r02d_test_model.log 2>&1; tail -12 gpurun_out/r02d_test_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r02d_ops_small.json > gpurun_out/r02d_bench_small.log 2>&1; tail -c 1800 gpurun_out/r02d_bench_small.log
