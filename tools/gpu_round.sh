#!/bin/bash
# One GPU-box session: full benches of the three configs, reference arm, ncu launch lists and --set full captures.
# Outputs under gpurun_out/ (copied to profiles/ by hand after reading them).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
for C in C3 C2 C5; do
  python bench.py --config $C > $O/r02_bench_$C.json 2> $O/r02_bench_$C.err
done
python bench.py --impl reference > $O/r02_ref_C3.json 2> $O/r02_ref_C3.err
NCU="ncu --clock-control none"
K='regex:k_shade|k_setup|k_raster|k_ssao|k_background|k_downsample|k_huge_prepare|k_emit|k_scan'
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r02_launches_C3.csv python bench.py --scenes 128 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r02_launches_C2.csv python bench.py --config C2 --scenes 128 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r02_launches_C5.csv python bench.py --config C5 --scenes 16 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
$NCU --set full --import-source on -k "$K" -c 40 -o $O/r02_prof_C3 -f python bench.py --scenes 64 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_prof_C3.log 2>&1
$NCU --set full --import-source on -k "$K" -c 40 -o $O/r02_prof_C2 -f python bench.py --config C2 --scenes 64 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_prof_C2.log 2>&1
$NCU --set full --import-source on -k "$K" -c 40 -o $O/r02_prof_C5 -f python bench.py --config C5 --scenes 8 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_prof_C5.log 2>&1
ls -la $O | tail -20
