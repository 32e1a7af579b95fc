mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -q -x -k "not baseline" 2>&1 | tail -2
timeout 300 python tools/parity_report.py small medium --batch 2 --out gpurun_out/r02p_parity.json > /dev/null 2>&1
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r02p_parity.json')):
    if 'error' in r: print(r); continue
    t2=r['T2']; print(r['config'], r['dtype'], 'logits rel %.2e abs %.2e box %.2e'%(max(v for k,v in t2.items() if k.endswith('logits_rel_l2')), max(v for k,v in t2.items() if k.endswith('logits_maxabs')), t2['boxes_maxabs']), 'mem %.2e'%r['T1']['memory'], 'blocks max %.2e'%max(v for k,v in r['T1'].items() if k.startswith('block')))
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-per-config --profile-out gpurun_out/r02p_ops_small.json > gpurun_out/r02p_bench_small.log 2>&1
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02p_ops_small.json')); print('sum_ms', r['sum_ms'])
for o in r['ops'][:7]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
l=[x for x in open('gpurun_out/r02p_bench_small.log') if x.startswith('{')][-1]
b=json.loads(l); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['with_fp32_host_input'])
PY
