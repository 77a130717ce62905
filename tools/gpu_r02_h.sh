#!/bin/bash
# Round 2, call H (1 GPU): final state of the round -- all GPU tests, smoke, bench, reference arm, kernel table.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "pytest -m gpu"
timeout 500 python -m pytest tests -m gpu -q -x --timeout 150 -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_gpu.log
el "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
el "bench"
timeout 420 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value %.4g (%.4f ms) nms %.4f link %.4f iou_frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"]["nms_frames_kernel"], d["roofline"]["kernels_ms"]["link_frames_kernel"], d["iou_matrix_roofline"]["frac"]))
    print("roofline", {k: d["roofline"][k] for k in ("frac", "achieved", "traffic", "traffic_source")}, d["roofline"].get("issue", {}).get("frac"))
    print("e2e %.4g (%.4f ms) host %s" % (e["value"], e["ms_per_step"], e["host_ms_per_step"]))
    print("ceiling", e["box_ceiling"]["upload_only_ms"], e["box_ceiling"]["stage_plus_upload_ms"], "registered", e["registered_inputs"]["ms_per_step"], "pinned", e["pinned_resubmit"]["ms_per_step"])
    print("parity", d["parity"])
    c = d["configs"]
    print("C3 link ms", c["config3_link"]["ms"], "C5", c["config5_video"]["kernels_ms"], "C4", c["config4_temporal"]["kernels_ms"])
    print("adapters", {k: (round(v["ms"], 2), round(v["cpu_ms"], 2), v["same"]) for k, v in d["adapters"]["calls"].items()})
except Exception as ex:
    print("bench unreadable", ex)
P
tail -n 3 gpurun_out/bench.err
el "reference arm"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_ref.json
el "kernel bench"
timeout 240 python tools/kernel_bench.py > gpurun_out/kernels.txt 2>> gpurun_out/h.err; cut -c1-150 gpurun_out/kernels.txt
el done
