#!/bin/bash
# Second evidence pass of round 1 (run under gpurun): the bench line after the e2e chunk change,
# the launch list of the bench command, ncu captures of the rewritten axis kernels and the
# surrounding-operation table.
O=gpurun_out
python bench.py > $O/r1b_bench.json 2> $O/r1b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1b_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/r1b_launch_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:dr_cols -s 4 -c 1 -o $O/cols_v2 python tools/surface_bench.py "sum(X, axis=0) float32" > $O/ncu_cols.log 2>&1
$NCU -k regex:dr_rows -s 4 -c 1 -o $O/rows_v2 python tools/surface_bench.py "sum(X, axis=1) float32" > $O/ncu_rows.log 2>&1
$NCU -k regex:dr_transpose -s 4 -c 1 -o $O/transpose_v1 python tools/surface_bench.py "X.T.copy() float32" > $O/ncu_tr.log 2>&1
python tools/surface_bench.py > $O/surface_final.log 2>&1
cut -c1-400 $O/r1b_bench.json; tail -3 $O/surface_final.log
