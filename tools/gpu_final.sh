# Final evidence run of a round: everything DESIGN.md / profiles/README.md cite, written under gpurun_out/ with the round prefix.
# .ncu-rep files are condensed to text ON the box (tools/profile_pack.py) and removed: gpurun only copies back 64 MiB.
R=${1:-r02}
mkdir -p gpurun_out
./tools/ubench/softmax_rate > gpurun_out/${R}_ubench_softmax.txt 2>&1
if [ "$2" = "tests" ]; then timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/${R}_test_all.log 2>&1; tail -3 gpurun_out/${R}_test_all.log; fi
timeout 400 python tools/bench_kernels.py --out gpurun_out/${R}_kernels_isolated.json > /dev/null 2>&1
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 sk=$3 ct=$4; shift 4
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $sk -c $ct -o gpurun_out/${R}_$name "$@" > /dev/null 2>&1
  python tools/profile_pack.py gpurun_out/${R}_$name.ncu-rep > gpurun_out/${R}_$name.txt 2>&1
  rm -f gpurun_out/${R}_$name.ncu-rep
}
cap ncu_attn_slots_small attn_slots 4 2 python tools/bench_kernels.py --only window_attention,global_attention --configs small --iters 2
cap ncu_attn_slots_medium attn_slots 4 2 python tools/bench_kernels.py --only window_attention,global_attention --configs medium --iters 2
cap ncu_msda_medium msda_fwd 3 1 python tools/bench_kernels.py --only msda_forward --configs medium --iters 3
cap ncu_msda_small msda_fwd 3 1 python tools/bench_kernels.py --only msda_forward --configs small --iters 3
cap ncu_gemm_small gemm_tc_kernel 5 4 python tools/one_forward.py --config small --batch 32 --n 1
cap ncu_conv3x3_small gemm_tc_kernel 42 2 python tools/one_forward.py --config small --batch 32 --n 1
# launch list of two eager forwards (shares of the step)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_ncu_launches_small_b32_fp16.csv python tools/one_forward.py --config small --batch 32 --n 2 > /dev/null 2>&1
# the bench line (with per_config) and the per-op table
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/${R}_ops_small.json > gpurun_out/${R}_bench_small.log 2>&1
grep '^{' gpurun_out/${R}_bench_small.log | tail -1 > gpurun_out/${R}_bench_small.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.log 2>&1
grep '^{' gpurun_out/${R}_bench_reference.log | tail -1 > gpurun_out/${R}_bench_reference.json
for c in medium; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-per-config --profile-out gpurun_out/${R}_ops_$c.json 2>/dev/null | grep '^{' | tail -1 > gpurun_out/${R}_bench_$c.json; done
rm -f gpurun_out/*.log
du -sh gpurun_out
python - <<PY
import json
b=json.load(open('gpurun_out/${R}_bench_small.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'])
PY
