#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -10
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; python -c "
import json
d=json.load(open('$OUT/bench_$TAG.json')); print('value %.0f ms/yr %.2f e2e %.0f frac %.3f traffic %s cpu %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value']))"
