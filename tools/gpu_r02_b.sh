#!/bin/bash
# Round 2, call B (1 GPU): parity tests + smoke + bench of the rebuilt e2e path, the rewritten big-frame NMS kernel.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 40 gpurun_out/pytest_gpu.log
el "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
el "big kernel time"
python tools/run_big_nms.py 296 | tee gpurun_out/big_time.txt
el "bench"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -n 8 gpurun_out/bench.err
el "h2d probe: sse vs avx512, thread sweep"
for th in 4 8 12 16; do
  VDET_HOST_COPY=avx512 timeout 120 python tools/h2d_scale_probe.py --threads $th --reps 30 >> gpurun_out/h2d_probe_avx512.json 2>> gpurun_out/b.err
done
VDET_HOST_COPY=sse timeout 120 python tools/h2d_scale_probe.py --threads 8 --reps 30 >> gpurun_out/h2d_probe_sse.json 2>> gpurun_out/b.err
python - <<'P'
import json
for f in ("gpurun_out/h2d_probe_avx512.json", "gpurun_out/h2d_probe_sse.json"):
    for line in open(f):
        d = json.loads(line)
        print(f.split("_")[-1], d["stage_threads"], {k: v["ms_per_round_max"] for k, v in d["legs"].items()})
P
el "ncu full: big kernel (rewritten)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_frames_big -s 2 -c 1 -o gpurun_out/prof_r02_big2 -f \
    python tools/run_big_nms.py 148 > gpurun_out/ncu_big2.log 2>&1; tail -n 2 gpurun_out/ncu_big2.log
el "kernel bench"
timeout 300 python tools/kernel_bench.py > gpurun_out/kernels.txt 2>> gpurun_out/b.err; grep -E "link|big|C5|C3" gpurun_out/kernels.txt
el done
