#!/bin/bash
# Round-2 profile capture on ONE GPU (run under gpurun): each ncu pass only after its own command exited 0 without ncu.
# Writes into gpurun_out/; tools/summarise_profiles.py turns the reports into the tracked files under profiles/.
set -x
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-extra"
$B > $O/r02_plain_bench.json 2> $O/r02_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $B > $O/r02_ncu_launches.log 2>&1
K="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra"
$K > $O/r02_plain_k3.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_run_lean -s 4 -c 1 -f -o $O/r02_k_run_lean_k3 $K > $O/r02_ncu_k3.log 2>&1
$K --n-ksub 24 > $O/r02_plain_k24.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_run_lean -s 4 -c 1 -f -o $O/r02_k_run_lean_k24 $K --n-ksub 24 > $O/r02_ncu_k24.log 2>&1
python tools/bench_ensemble.py --particles 128 > $O/r02_plain_ens.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ensemble -c 1 -f -o $O/r02_k_ensemble python tools/bench_ensemble.py --particles 128 > $O/r02_ncu_ens.log 2>&1
ls -la $O/*.ncu-rep
