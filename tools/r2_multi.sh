#!/bin/bash
# one 8-GPU box: sharded parity checks, strong scaling at n=30 (N=2,4,8), weak n=33 on 8 GPUs
# usage: tools/r2_multi.sh [tag] [list of N]   (default tag r2, N = 2 4 8)
mkdir -p gpurun_out
TAG=${1:-r2}; shift
NS=${@:-2 4 8}
P=29540
for N in $NS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+N)) tests/dist_sharded_check.py 20 2> gpurun_out/${TAG}_sharded_check_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_sharded_check_${N}gpu.json; echo "check N=$N rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+10+N)) bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench_strong_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_bench_strong_${N}gpu.json; echo "bench N=$N rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_strong_${N}gpu.json')); print('N=$N', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['engine'].get('peer_pass_ms_mean'), d['engine']['passes_per_layer'], d['check']['ok'], d['check']['at_bench_size'])"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+30)) bench.py --gpus 8 --weak --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench_n33_weak_8gpu.err | grep '^{' > gpurun_out/${TAG}_bench_n33_weak_8gpu.json; echo "weak rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n33_weak_8gpu.json')); print('n33', d['value'], 'e2e', d['e2e']['value'], json.dumps(d['check']))"
