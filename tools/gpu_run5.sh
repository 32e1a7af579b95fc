mkdir -p gpurun_out
timeout 120 python tools/attn_debug.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r02e_test_attn.log 2>&1; tail -5 gpurun_out/r02e_test_attn.log
timeout 300 python tools/bench_kernels.py --only window_attention,global_attention --configs small,medium --out gpurun_out/r02e_kernels_attn.json 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*' | grep -v msda
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 2 -c 2 -o gpurun_out/r02e_ncu_attn_small python tools/bench_kernels.py --only window_attention,global_attention --configs small --iters 1 > gpurun_out/r02e_ncu_attn.log 2>&1; tail -2 gpurun_out/r02e_ncu_attn.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_gemm_gpu.py -x -q -k "not baseline and not medium and not xlarge and not large" > gpurun_out/r02e_test_model.log 2>&1; tail -12 gpurun_out/r02e_test_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r02e_ops_small.json > gpurun_out/r02e_bench_small.log 2>&1; tail -c 600 gpurun_out/r02e_bench_small.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02e_ops_small.json')); print('sum_ms', r['sum_ms'])
for o in r['ops'][:14]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
PY
