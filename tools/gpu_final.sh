#!/bin/bash
# Short confirmation round: GPU parity tests, smoke, bench, then the two ncu passes (killed once captured).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 150 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
el smoke
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
el bench
timeout 120 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-300; tail -n 3 gpurun_out/bench.err
el "ncu full"
VDET_E2E_MODE=sync timeout 90 ncu --set full --clock-control none --import-source on -k regex:'nms_frames_kernel|link_frames_kernel' \
    -s 4 -c 3 --kill 1 -o gpurun_out/prof_r01 -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
el "ncu launch list"
VDET_E2E_MODE=sync timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --kill 1 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | head -20
el done
