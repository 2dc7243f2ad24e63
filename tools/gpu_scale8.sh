#!/bin/bash
# 8-GPU check of the end-to-end leg: one group per GPU (round-1 form) against two overlapped groups.  gpurun --gpus 8
TAG=${1:-s8}
OUT=gpurun_out
mkdir -p $OUT
N=${2:-8}
for G in 1 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --e2e-groups $G \
     > $OUT/bench_n${N}_g${G}_$TAG.json 2> $OUT/bench_n${N}_g${G}_$TAG.err
  python -c "
import json
d=json.load(open('$OUT/bench_n${N}_g${G}_$TAG.json')); print('N=$N groups=$G: value %.3fM ms/yr %.2f e2e %.3fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))" || tail -5 $OUT/bench_n${N}_g${G}_$TAG.err
done | tee $OUT/scale_$TAG.log
