mkdir -p gpurun_out
for m in clean dirty; do timeout 300 python tools/bench_kernels.py --only msda_forward --flush $m --out gpurun_out/r02b_kernels_msda_$m.json 2>&1 | cut -c1-200 | grep -o '"config": "[^"]*", "us_median": [0-9.]*\|frac": [0-9.]*' | paste - - ; done
timeout 900 python tools/parity_report.py tiny small medium large xlarge --batch 2 --out gpurun_out/r02b_parity_ladder.json > gpurun_out/r02b_parity_ladder.log 2>&1
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r02b_parity_ladder.json')):
    if 'error' in r: print(r); continue
    t2=r['T2']; print(r['config'], r['dtype'], 'rel %.2e abs %.2e box %.2e auxbox %.2e encbox %.2e'%(max(v for k,v in t2.items() if k.endswith('logits_rel_l2')), max(v for k,v in t2.items() if k.endswith('logits_maxabs')), t2['boxes_maxabs'], max(t2['aux0_boxes_maxabs'],t2['aux1_boxes_maxabs']), t2['enc_boxes_maxabs']), 'mem %.2e score %.2e'%(r['T1']['memory'], r['T1']['enc_score_maxabs']))
PY
timeout 1500 python -m pytest tests/test_model_gpu.py -q -k "baseline" > gpurun_out/r02b_test_baseline.log 2>&1; tail -15 gpurun_out/r02b_test_baseline.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/parity_baseline_*.json')):
    r=json.load(open(f)); print({k:(round(v,6) if isinstance(v,float) else v) for k,v in r.items()})
PY
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_dropin_demo.py -q -x > gpurun_out/r02b_test_kernels.log 2>&1; tail -8 gpurun_out/r02b_test_kernels.log
