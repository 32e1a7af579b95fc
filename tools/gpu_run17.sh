mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -q -x -k "not baseline" 2>&1 | tail -2
LWDETR_BENCH_E2E_AB=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-per-config --profile-out gpurun_out/r02o_ops_small.json > gpurun_out/r02o_bench_small.log 2> gpurun_out/r02o_bench_small.err; grep "A/B" gpurun_out/r02o_bench_small.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02o_ops_small.json')); print('sum_ms', r['sum_ms'])
for o in r['ops'][:14]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
l=[x for x in open('gpurun_out/r02o_bench_small.log') if x.startswith('{')][-1]
b=json.loads(l); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['with_fp32_host_input'])
PY
LWDETR_B200_GEMM_GROUPS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-per-config --profile-out gpurun_out/r02o_ops_small_g1.json > gpurun_out/r02o_bench_small_g1.log 2>&1
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02o_ops_small_g1.json')); print('GROUPS=1 sum_ms', r['sum_ms'])
for o in r['ops'][:8]: print('%-16s n=%2d %8.1f us share %.3f'%(o['op'],o['launches'],o['ms']*1e3,o['share']))
l=[x for x in open('gpurun_out/r02o_bench_small_g1.log') if x.startswith('{')][-1]
b=json.loads(l); print(b['value'], b['ms_per_step'])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 4 -c 1 -o gpurun_out/r02o_ncu_win_small python tools/bench_kernels.py --only window_attention --configs small --iters 2 > /dev/null 2>&1
