mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -q -x -k "not baseline" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-per-config --profile-out gpurun_out/r02m_ops_small.json > gpurun_out/r02m_bench_small.log 2>&1
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02m_ops_small.json')); print('sum_ms', r['sum_ms'])
for o in r['ops'][:14]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
l=[x for x in open('gpurun_out/r02m_bench_small.log') if x.startswith('{')][-1]
b=json.loads(l); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['with_fp32_host_input'])
PY
