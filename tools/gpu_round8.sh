#!/bin/bash
# One 8-GPU box session: raw D2H ceiling at 1/2/4/8 processes, the C5 1 -> 8 sweep, C3 at 8 GPUs (device-resident + e2e).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-r02}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/d2h_probe.py > $O/${T}_d2h_probe_n1.json 2> $O/${T}_d2h_probe.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29600 + N)) tools/d2h_probe.py > $O/${T}_d2h_probe_n$N.json 2>> $O/${T}_d2h_probe.err
done
cat $O/${T}_d2h_probe_n*.json
python bench.py --config C5 --no-cpu > $O/${T}_bench_C5_n1.json 2> $O/${T}_bench_C5_n1.err
for N in 2 4 8; do
  $TR --nproc-per-node $N --master-port $((29700 + N)) bench.py --config C5 --gpus $N --no-cpu > $O/${T}_bench_C5_n$N.json 2> $O/${T}_bench_C5_n$N.err
done
$TR --nproc-per-node 8 --master-port 29808 bench.py --gpus 8 --no-cpu > $O/${T}_bench_C3_n8.json 2> $O/${T}_bench_C3_n8.err
for f in $O/${T}_bench_C5_n*.json $O/${T}_bench_C3_n8.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d["n_gpus"], "gpus", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "link", d["e2e"].get("d2h_link_gbs"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
