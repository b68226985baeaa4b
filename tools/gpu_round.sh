#!/bin/bash
# One GPU-box session: GPU tests, full benches of the three configs, reference arm, ncu launch lists and --set full captures.
# Outputs under gpurun_out/ (copied to profiles/ by hand after reading them). gpurun pulls at most 64 MiB back, so the
# C2 / C5 reports are reduced to tools/ncu_metrics.py numbers on the box and deleted; only the C3 report travels.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-r02}
python -m pytest tests -q -m gpu -x > $O/${T}_gputests.log 2>&1; tail -3 $O/${T}_gputests.log
for C in C3 C2 C5; do
  python bench.py --config $C > $O/${T}_bench_$C.json 2> $O/${T}_bench_$C.err
done
python bench.py --impl reference > $O/${T}_ref_C3.json 2> $O/${T}_ref_C3.err
NCU="ncu --clock-control none"
K='regex:k_shade|k_setup|k_raster|k_ssao|k_background|k_downsample|k_huge_prepare|k_emit|k_scan'
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${T}_launches_C3.csv python bench.py --scenes 128 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${T}_launches_C2.csv python bench.py --config C2 --scenes 128 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${T}_launches_C5.csv python bench.py --config C5 --scenes 16 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
$NCU --set full --import-source on -k "$K" -c 14 -o $O/${T}_prof_C3 -f python bench.py --scenes 64 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/${T}_prof_C3.log 2>&1
$NCU --set full -k "$K" -c 24 -o /tmp/prof_C2 -f python bench.py --config C2 --scenes 64 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/${T}_prof_C2.log 2>&1
$NCU --set full -k "$K" -c 24 -o /tmp/prof_C5 -f python bench.py --config C5 --scenes 8 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/${T}_prof_C5.log 2>&1
python tools/aux_kernels.py > $O/${T}_aux_timing.json 2> $O/${T}_aux.err
$NCU --set full -k 'regex:k_cam|k_png|k_jpeg|k_pose' -c 40 -o /tmp/prof_aux -f python tools/aux_kernels.py --no-timing > $O/${T}_prof_aux.log 2>&1
python tools/ncu_summary.py full /tmp/prof_aux.ncu-rep > $O/${T}_aux_kernels.md 2>> $O/${T}_aux.err
python tools/ncu_metrics.py C3=$O/${T}_prof_C3.ncu-rep:64 C2=/tmp/prof_C2.ncu-rep:64 C5=/tmp/prof_C5.ncu-rep:8 > $O/${T}_ncu_metrics.log 2>&1
cp profiles/r02_ncu_metrics.json $O/${T}_ncu_metrics.json
ls -la $O | tail -30
