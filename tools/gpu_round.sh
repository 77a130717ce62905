#!/bin/bash
# One gpurun call = smoke + GPU parity tests + bench (+ optional ncu passes).  Everything lands in gpurun_out/.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [ncu]'
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" | tee gpurun_out/smoke.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 40 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
if [ "${1:-}" = "ncu" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  echo "== ncu full: nms_frames + link + iou_matrix"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nms_frames_kernel|link_frames_kernel|iou_matrix_f32' \
      -s 6 -c 5 -o gpurun_out/prof_r01 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'iou_matrix_f32' -c 1 \
      -o gpurun_out/prof_iou_r01 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_iou.log 2>&1
  ls -la gpurun_out
fi
if [ "${1:-}" = "multi" ]; then
  N=$(nvidia-smi -L | wc -l)
  echo "== multi-GPU parity on $N GPUs"
  timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider 2>&1 | tail -n 15 | tee gpurun_out/pytest_multi.log
  for n in 2 4 8; do
    if [ $n -le $N ]; then
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
          bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
      echo "bench n=$n rc=$?"; cat gpurun_out/bench_n$n.json; tail -n 3 gpurun_out/bench_n$n.err
    fi
  done
fi
