mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_test_all.log 2>&1; tail -5 gpurun_out/r02k_test_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r02k_ops_small.json > gpurun_out/r02k_bench_small.log 2>&1; tail -c 1500 gpurun_out/r02k_bench_small.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02k_ops_small.json')); print('sum_ms', r['sum_ms'])
for o in r['ops'][:12]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
PY
