mkdir -p gpurun_out
./tools/ubench/stream_rate > gpurun_out/r02c_ubench_stream.txt 2>&1; cat gpurun_out/r02c_ubench_stream.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r02c_test_attn.log 2>&1; tail -25 gpurun_out/r02c_test_attn.log
timeout 300 python tools/bench_kernels.py --only window_attention,global_attention --out gpurun_out/r02c_kernels_attn.json 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*' 
LWDETR_B200_ATTN_SLOTS=0 timeout 300 python tools/bench_kernels.py --only window_attention,global_attention --configs small,medium 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*' | sed 's/^/OLD /'
LWDETR_B200_ATTN_SLOTS=2 timeout 300 python tools/bench_kernels.py --only global_attention --configs medium,large 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*' | sed 's/^/SLOTS2 /'
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "not baseline and not medium and not xlarge and not large" > gpurun_out/r02c_test_model.log 2>&1; tail -25 gpurun_out/r02c_test_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r02c_ops_small.json > gpurun_out/r02c_bench_small.log 2>&1; tail -c 2500 gpurun_out/r02c_bench_small.log
