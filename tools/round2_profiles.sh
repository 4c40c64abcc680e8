#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun): the default bench line, its launch list, the
# sustained Black-Scholes figure, one `ncu --set full` capture of the Black-Scholes kernel and of
# the halo-pushing stencil kernel, and the n-body contraction-share experiment.
set -x
O=gpurun_out
python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err
python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu --no-others > $O/r2_bench_bs_200steps.json 2>> $O/r2_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > $O/r2_launch_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:dr_flat -s 2 -c 1 -o $O/r2_bs python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-others > $O/r2_ncu1.log 2>&1
$NCU -k regex:dr_stencil -s 12 -c 1 -o $O/r2_stencil_halo python tools/heat_shard_probe.py 16384 2 > $O/r2_ncu2.log 2>&1
$NCU -k regex:dr_stencil -s 3 -c 1 -o $O/r2_stencil_base python tools/heat_shard_probe.py 16384 1 > $O/r2_ncu3.log 2>&1
python tools/nbody_contraction_share.py > $O/r2_nbody_contraction_share.txt 2>&1
cut -c1-400 $O/r2_bench_default.json; cut -c1-300 $O/r2_bench_bs_200steps.json; cat $O/r2_nbody_contraction_share.txt; ls -la $O/*.ncu-rep
