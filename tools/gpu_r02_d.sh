#!/bin/bash
# Round 2, call D (1 GPU): bench line + ncu launch list + --set full captures of the step's kernels.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "targeted tests"
timeout 200 python -m pytest tests/test_gpu_nms.py -m gpu -q -x --timeout 100 -p no:cacheprovider -k "postprocessor or compact or big or frames_vs" > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_d.log
el "bench"
timeout 420 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value %.4g (%.4f ms) nms %.4f link %.4f iou_frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"]["nms_frames_kernel"], d["roofline"]["kernels_ms"]["link_frames_kernel"], d["iou_matrix_roofline"]["frac"]))
    print("e2e %.4g (%.4f ms) wall %.4f host %s" % (e["value"], e["ms_per_step"], e["host_wall_ms_per_step"], e["host_ms_per_step"]))
    print("registered", e["registered_inputs"]["ms_per_step"], e["registered_inputs"]["consistent"], "pinned", e["pinned_resubmit"]["ms_per_step"])
    print("parity", d["parity"])
    print("configs", json.dumps(d.get("configs"))[:2500])
    print("adapters", json.dumps(d.get("adapters"))[:2500])
except Exception as ex:
    print("bench unreadable", ex)
P
tail -n 5 gpurun_out/bench.err
el "ncu launch list (bench, extras off)"
VDET_BENCH_EXTRAS=0 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
el "ncu full: step kernels"
VDET_BENCH_EXTRAS=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:'nms_frames_kernel|link_frames_kernel|iou_matrix_f32|compact_keep|keep_offsets' \
    -s 8 -c 6 -o gpurun_out/prof_r02 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"; tail -n 2 gpurun_out/ncu_full.log
el done
