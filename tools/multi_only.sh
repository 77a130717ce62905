#!/bin/bash
# Multi-GPU verification only (no single-GPU suite): parity test on all GPUs + bench at N = 2,4,8 + host/device probe.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
N=$(nvidia-smi -L | wc -l)
echo "== multi-GPU parity on $N GPUs"
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider 2>&1 | tail -n 15 | tee gpurun_out/pytest_multi.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --impl reference --gpus $n --steps 3 --warmup 1 > gpurun_out/bench_ref_n$n.json 2> gpurun_out/bench_ref_n$n.err
    echo "ref n=$n rc=$?"; head -c 300 gpurun_out/bench_ref_n$n.json; echo
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 \
        bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    echo "bench n=$n rc=$?"; cat gpurun_out/bench_n$n.json | cut -c1-400; tail -n 3 gpurun_out/bench_n$n.err
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 \
        tools/multi_probe.py > gpurun_out/multi_probe_n$n.json 2> gpurun_out/multi_probe_n$n.err
    echo "probe n=$n rc=$?"; cat gpurun_out/multi_probe_n$n.json | tr -d '\n'; echo
  fi
done
