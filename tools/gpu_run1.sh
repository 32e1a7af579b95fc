mkdir -p gpurun_out
./tools/ubench/softmax_rate > gpurun_out/r02a_ubench_softmax.txt 2>&1
nproc > gpurun_out/r02a_host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r02a_host.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "msda or ms_deform" > gpurun_out/r02a_test_msda.log 2>&1; tail -5 gpurun_out/r02a_test_msda.log
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "tiny or large-1" > gpurun_out/r02a_test_model.log 2>&1; tail -5 gpurun_out/r02a_test_model.log
timeout 300 python tools/bench_kernels.py --only msda_forward --out gpurun_out/r02a_kernels_msda.json > gpurun_out/r02a_kernels_msda.log 2>&1; cat gpurun_out/r02a_kernels_msda.log | cut -c1-220
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench_small.log 2>&1; tail -c 1500 gpurun_out/r02a_bench_small.log
cat gpurun_out/r02a_ubench_softmax.txt gpurun_out/r02a_host.txt
