#!/bin/bash
# One GPU-box session: GPU parity tests, the default bench line, the ncu launch list of the same bench command and
# one `ncu --set full` capture of the two tstepo kernels.  Run as:  gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r1c'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/nproc_$TAG.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/smi_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
cat $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
cat $OUT/bench_ref_$TAG.json
# launch list of the same command (first launches: build + spin-up year; kernel SHARES are what is compared)
# (bench.py spins the ocean up for 100 untimed model years = 249 600 launches, + 2496 of the warm-up year: skipped)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 252200 -c 1500 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
# full capture of the tracer / barotropic / BIOGEM kernels at the bench's member count, 100-year-old state
# (9600 ocean cycles x 5.5 matching launches per cycle are skipped)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tstep_col|k_co_col|k_baro_blk|k_bg_step|k_tc_partial|k_tc_apply" -s 52800 -c 14 \
  -o $OUT/prof_tstepo_$TAG -f python tools/prof_run.py --members 128 --spin 9600 --steps 8 --variant col > $OUT/prof_full_$TAG.log 2>&1
ncu -i $OUT/prof_tstepo_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_tstepo_$TAG.ncu-rep --page details --csv > $OUT/details_$TAG.csv 2>/dev/null
timeout 300 python tools/trace_run.py > $OUT/trace_$TAG.log 2>&1
ls -la $OUT | tail -12
