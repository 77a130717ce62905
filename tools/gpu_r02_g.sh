#!/bin/bash
# Round 2, call G (1 GPU): the x-sorted link -- parity, A/B against the full scan, bench.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "link parity tests"
timeout 300 python -m pytest tests/test_gpu_iou_link.py tests/test_gpu_full_configs.py tests/test_gpu_multi.py -m gpu -q -x --timeout 120 -p no:cacheprovider > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_g.log
el "A/B: sorted link vs full scan (kernel bench rows)"
timeout 200 python tools/kernel_bench.py > gpurun_out/kernels_sorted.txt 2>> gpurun_out/g.err; grep -E "link" gpurun_out/kernels_sorted.txt | cut -c1-150
VDET_LINK_NO_SORT=1 timeout 200 python tools/kernel_bench.py > gpurun_out/kernels_nosort.txt 2>> gpurun_out/g.err; grep -E "link" gpurun_out/kernels_nosort.txt | cut -c1-150
el "full pytest"
timeout 500 python -m pytest tests -m gpu -q -x --timeout 150 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
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
    print("e2e %.4g (%.4f ms) host %s" % (e["value"], e["ms_per_step"], e["host_ms_per_step"]))
    print("ceiling", e["box_ceiling"]["upload_only_ms"], e["box_ceiling"]["stage_plus_upload_ms"], "registered", e["registered_inputs"]["ms_per_step"], "pinned", e["pinned_resubmit"]["ms_per_step"])
    print("parity", d["parity"])
    c = d["configs"]
    print("C3 link ms", c["config3_link"]["ms"], "C5", c["config5_video"]["kernels_ms"], "C4", c["config4_temporal"]["kernels_ms"])
except Exception as ex:
    print("bench unreadable", ex)
P
tail -n 3 gpurun_out/bench.err
el "ncu: link kernels"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'link_frames|sort_frames' -s 4 -c 4 -o gpurun_out/prof_r02_link -f \
    python tools/kernel_bench.py --quick > gpurun_out/ncu_link.log 2>&1; tail -n 1 gpurun_out/ncu_link.log
el done
