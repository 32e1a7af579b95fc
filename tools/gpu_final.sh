# Final evidence run of a round: everything DESIGN.md / profiles/README.md cite, written under gpurun_out/ with the round prefix.
R=${1:-r02}
mkdir -p gpurun_out
./tools/ubench/softmax_rate > gpurun_out/${R}_ubench_softmax.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/${R}_test_all.log 2>&1; tail -3 gpurun_out/${R}_test_all.log
timeout 400 python tools/bench_kernels.py --out gpurun_out/${R}_kernels_isolated.json > gpurun_out/${R}_kernels_isolated.log 2>&1
# ncu --set full captures (one launch each, isolated kernels at the BASELINE shapes)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 4 -c 2 -o gpurun_out/${R}_ncu_attn_slots_small python tools/bench_kernels.py --only window_attention,global_attention --configs small --iters 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_slots -s 4 -c 2 -o gpurun_out/${R}_ncu_attn_slots_medium python tools/bench_kernels.py --only window_attention,global_attention --configs medium --iters 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 1 -o gpurun_out/${R}_ncu_msda_medium python tools/bench_kernels.py --only msda_forward --configs medium --iters 3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 3 -c 1 -o gpurun_out/${R}_ncu_msda_small python tools/bench_kernels.py --only msda_forward --configs small --iters 3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 4 -o gpurun_out/${R}_ncu_gemm_small python tools/one_forward.py --config small --batch 32 --n 1 > /dev/null 2>&1
# launch list of two eager forwards (shares of the step)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_ncu_launches_small_b32_fp16.csv python tools/one_forward.py --config small --batch 32 --n 2 > /dev/null 2>&1
# the bench line (with per_config) and the per-op table
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/${R}_ops_small.json > gpurun_out/${R}_bench_small.log 2>&1
grep '^{' gpurun_out/${R}_bench_small.log | tail -1 > gpurun_out/${R}_bench_small.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.log 2>&1
grep '^{' gpurun_out/${R}_bench_reference.log | tail -1 > gpurun_out/${R}_bench_reference.json
for c in medium large; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-per-config --profile-out gpurun_out/${R}_ops_$c.json 2>/dev/null | grep '^{' | tail -1 > gpurun_out/${R}_bench_$c.json; done
python - <<PY
import json
b=json.load(open('gpurun_out/${R}_bench_small.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'])
PY
