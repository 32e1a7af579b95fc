mkdir -p gpurun_out
G='"kernel": "[a-z_]*", "config": "[^"]*", "us_median": [0-9.]*'
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r02l_test_all.log 2>&1; tail -5 gpurun_out/r02l_test_all.log
timeout 300 python tools/bench_kernels.py --out gpurun_out/r02l_kernels.json 2>&1 | grep -o "$G"
timeout 600 python bench.py --steps 20 --warmup 5 --no-per-config --profile-out gpurun_out/r02l_ops_small.json > gpurun_out/r02l_bench_small.log 2>&1; tail -c 600 gpurun_out/r02l_bench_small.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02l_ops_small.json')); print('sum_ms', r['sum_ms'])
for o in r['ops'][:12]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
l=[x for x in open('gpurun_out/r02l_bench_small.log') if x.startswith('{')][-1]
b=json.loads(l); print(b['value'], b['ms_per_step'], b['e2e']['value'])
PY
