mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02f_test_all.log 2>&1; tail -30 gpurun_out/r02f_test_all.log
timeout 400 python tools/bench_kernels.py --out gpurun_out/r02f_kernels.json 2>&1 | grep -o '"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*'
timeout 300 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r02f_ops_small.json > gpurun_out/r02f_bench_small.log 2>&1; tail -c 2500 gpurun_out/r02f_bench_small.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02f_ops_small.json')); print('sum_ms', r['sum_ms'])
for o in r['ops'][:24]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
PY
