#!/bin/bash
# One gpurun call: smoke + GPU parity tests + bench + e2e variants (+ ncu).  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh [ncu]'
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
el "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
el "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
el "bench"
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
el "bench reference arm"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
cat gpurun_out/bench_ref.json
el "e2e variants"
timeout 300 python tools/e2e_variants.py 100 > gpurun_out/e2e_variants.json 2> gpurun_out/e2e_variants.err; echo "variants rc=$?"
cat gpurun_out/e2e_variants.json | tr -d '\n' | head -c 6000; echo; tail -n 3 gpurun_out/e2e_variants.err
if [ "${1:-}" = "ncu" ]; then
  el "ncu launch list"
  VDET_E2E_MODE=sync timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  el "ncu full: nms_frames + link + iou_matrix"
  VDET_E2E_MODE=sync timeout 500 ncu --set full --clock-control none --import-source on -k regex:'nms_frames_kernel|link_frames_kernel|iou_matrix_f32' \
      -s 6 -c 5 -o gpurun_out/prof_r01 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
el "done"
