#!/bin/bash
# Round 2, call E (1 GPU): two-array sort variants -- parity, A/B timing against the single-array network, bench.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "NMS parity tests"
timeout 400 python -m pytest tests/test_gpu_nms.py tests/test_gpu_packed.py tests/test_gpu_full_configs.py -m gpu -q -x --timeout 150 -p no:cacheprovider > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_e.log
el "A/B: split vs single-array network"
rm -f gpurun_out/nms_time_split.json gpurun_out/nms_time_nosplit.json
for i in 1 2 3; do
timeout 60 python tools/nms_time.py 1000 300 30 50 >> gpurun_out/nms_time_split.json 2>> gpurun_out/e.err
VDET_NMS_NO_SPLIT=1 timeout 60 python tools/nms_time.py 1000 300 30 50 >> gpurun_out/nms_time_nosplit.json 2>> gpurun_out/e.err
done
cat gpurun_out/nms_time_split.json gpurun_out/nms_time_nosplit.json
for n in 150 190 380; do
timeout 60 python tools/nms_time.py 1000 $n 30 30 >> gpurun_out/nms_time_split_sizes.json 2>> gpurun_out/e.err
VDET_NMS_NO_SPLIT=1 timeout 60 python tools/nms_time.py 1000 $n 30 30 >> gpurun_out/nms_time_split_sizes.json 2>> gpurun_out/e.err
done
cat gpurun_out/nms_time_split_sizes.json
el "full pytest"
timeout 400 python -m pytest tests -m gpu -q -x --timeout 150 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
el "bench"
VDET_BENCH_EXTRAS=0 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_e.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench_e.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value %.4g (%.4f ms) nms %.4f link %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"]["nms_frames_kernel"], d["roofline"]["kernels_ms"]["link_frames_kernel"]))
    print("e2e %.4g (%.4f ms) host %s registered %.4f pinned %.4f" % (e["value"], e["ms_per_step"], e["host_ms_per_step"], e["registered_inputs"]["ms_per_step"], e["pinned_resubmit"]["ms_per_step"]))
except Exception as ex:
    print("bench unreadable", ex)
P
tail -n 3 gpurun_out/bench.err
el done
