#!/bin/bash
# Round-2 profiling pass (one B200 through gpurun): launch list + one `ncu --set full` capture per kernel of the step and of the
# stand-alone MLP kernel. Summaries are extracted HERE afterwards with scripts/ncu_summary.py and committed under profiles/.
P=gpurun_out/prof_r2
mkdir -p $P
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -s 20 -c 40 --csv --log-file $P/launches.csv python scripts/profile_step.py 12 > $P/launches.log 2>&1
$NCU --set full --import-source on -k regex:fused_ws -s 3 -c 1 -o $P/prof_fused python scripts/profile_step.py 5 > $P/prof_fused.log 2>&1
$NCU --set full --import-source on -k regex:adam_step -s 3 -c 1 -o $P/prof_adam python scripts/profile_step.py 5 > $P/prof_adam.log 2>&1
$NCU --set full --import-source on -k regex:bin_ -s 9 -c 3 -o $P/prof_binning python scripts/profile_step.py 5 > $P/prof_binning.log 2>&1
$NCU --set full --import-source on -k regex:mlp_forward -s 2 -c 1 -o $P/prof_mlp128x8 python scripts/bench_mlp.py --widths 128 --hidden 8 --batches 1048576 --iters 3 > $P/prof_mlp128x8.log 2>&1
$NCU --set full --import-source on -k regex:mlp_forward -s 2 -c 1 -o $P/prof_mlp128x4 python scripts/bench_mlp.py --widths 128 --hidden 4 --batches 1048576 --iters 3 > $P/prof_mlp128x4.log 2>&1
$NCU --set full --import-source on -k regex:mlp_forward -s 2 -c 1 -o $P/prof_mlp64x4 python scripts/bench_mlp.py --widths 64 --hidden 4 --batches 1048576 --iters 3 > $P/prof_mlp64x4.log 2>&1
ls -la $P
