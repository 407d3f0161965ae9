#!/bin/bash
# Evidence capture for one round (run on the GPU box through gpurun):  bash profiles/capture.sh r2
# Writes into gpurun_out/ (scratch); profiles/summarize.py turns the reports into the committed CSV/JSON summaries.
# Numbers printed by processes running under ncu are never bench values.
# CAPTURE_PARTS=ba limits the --set full captures to the back-end kernels (the front end has not changed since the last capture).
R=${1:-r2}
PARTS=${CAPTURE_PARTS:-all}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list (per-launch durations) of the bench command itself (configs[1], the headline workload)
$NCU --metrics gpu__time_duration.sum -c 12000 --csv --log-file gpurun_out/${R}_bench_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/${R}_bench_under_ncu.log 2>&1
# 2. --set full captures of single launches in steady state (window full, marginalisation running), one sequence
FULL="$NCU --set full --import-source on -f"
FULLNS="$NCU --set full -f"
$FULL -k regex:marg_solve -s 4 -c 1 -o gpurun_out/${R}_marg_solve python harness/debug_backend.py 16 > gpurun_out/${R}_ncu_marg.log 2>&1
$FULL -k regex:ba_step -s 20 -c 1 -o gpurun_out/${R}_ba_step python harness/debug_backend.py 16 > gpurun_out/${R}_ncu_step.log 2>&1
$FULLNS -k "regex:ba_eval|ba_reduce|marg_eval|marg_gather|preint_jobs|ba_finish" -s 40 -c 10 -o gpurun_out/${R}_ba_small python harness/debug_backend.py 16 > gpurun_out/${R}_ncu_small.log 2>&1
if [ "$PARTS" = all ]; then
$FULL -k regex:lk_track -s 12 -c 1 -o gpurun_out/${R}_lk_track python harness/run_tracker.py --frames 24 > gpurun_out/${R}_ncu_lk.log 2>&1
$FULLNS -k "regex:clahe|pyrdown|min_eig|gftt|sort_keys|mask_" -s 60 -c 10 -o gpurun_out/${R}_fe_small python harness/run_tracker.py --frames 24 > gpurun_out/${R}_ncu_fe.log 2>&1
fi
# 3. the same kernels serving a batch of 64 sequences (BASELINE configs[2]): one launch per stage for all members
[ "$PARTS" = all ] && $FULLNS -k "regex:lk_track|min_eig|clahe_apply|pyrdown" -s 150 -c 6 -o gpurun_out/${R}_batch_fe python harness/run_batch.py --seqs 64 --steps 3 > gpurun_out/${R}_ncu_batch_fe.log 2>&1
$FULLNS -k "regex:ba_eval|ba_reduce|ba_step|marg_solve" -s 60 -c 6 -o gpurun_out/${R}_batch_ba python harness/run_batch.py --seqs 64 --steps 3 > gpurun_out/${R}_ncu_batch_ba.log 2>&1
# raw pages as CSV (what profiles/summarize.py reads); only the three single-launch reports with source are kept as .ncu-rep
for f in gpurun_out/${R}_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
for f in ba_small fe_small batch_fe batch_ba; do rm -f gpurun_out/${R}_$f.ncu-rep; done
# 4. microbenchmarks that sized the single-CTA solvers (cycle counters, not wall clock)
[ "$PARTS" = all ] && for b in lat chol_bench eig_bench; do [ -x harness/micro/$b ] && harness/micro/$b > gpurun_out/${R}_micro_$b.txt 2>&1; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > gpurun_out/${R}_clocks.csv
ls -la gpurun_out/ | tail -20
