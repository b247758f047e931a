#!/bin/bash
# round 2, call F (N GPUs): distributed parity test + bench.py --gpus N exactly as the driver launches it
N=${1:-2}
STEPS=${2:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
if [ "$N" = "2" ]; then
  step "2-GPU parity (tests/test_distributed.py)"
  timeout 600 python -m pytest tests/test_distributed.py -q -m gpu --timeout 500 -x > $O/f${N}_tests.log 2>&1; echo "exit=$?" >> $O/f${N}_tests.log; tail -8 $O/f${N}_tests.log
fi
step "bench N=$N, 100 M reads per GPU"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 1 --no-cpu-baseline > $O/f${N}_bench100.json 2> $O/f${N}_bench100.err
echo "rc=$?"; tail -6 $O/f${N}_bench100.err; tail -c 2600 $O/f${N}_bench100.json
step "done"
