#!/bin/bash
# round 2, call C (2 GPUs): the 2-GPU parity test, then bench.py --gpus 2 as the driver launches it (20 M and 100 M reads per GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
step() { echo "== $1 (t=$(( $(date +%s) - T0 ))s)"; }
nvidia-smi topo -m > $O/c_topo.txt 2>&1
step "2-GPU parity (tests/test_distributed.py)"
timeout 600 python -m pytest tests/test_distributed.py -q -m gpu --timeout 500 -x > $O/c_tests.log 2>&1; echo "exit=$?" >> $O/c_tests.log; tail -25 $O/c_tests.log
for R in 20000000 100000000; do
  step "bench N=2 reads/GPU=$R"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --reads $R --no-cpu-baseline > $O/c_bench2_$R.json 2> $O/c_bench2_$R.err
  echo "rc=$?"; tail -5 $O/c_bench2_$R.err; tail -c 2500 $O/c_bench2_$R.json
done
step "done"
