# usage: bash tools/gpu_multi.sh <ngpus> <config>   (inside gpurun --gpus N)
N=$1; C=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $C --steps 20 --warmup 5 > gpurun_out/r02_bench_${C}_n${N}.log 2>&1
grep '^{' gpurun_out/r02_bench_${C}_n${N}.log | tail -1 > gpurun_out/r02_bench_${C}_n${N}.json
python - <<PY
import json
b=json.load(open('gpurun_out/r02_bench_${C}_n${N}.json')); print('$C', 'N=$N', b['value'], b['ms_per_step'], b['e2e']['value'], b['config'].get('replicas_bit_identical'), b['p50_latency_bs1_ms'])
PY
