#!/bin/bash
# kernel experiment round: parity tests first, then NMS timings under the residency caps
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
: > gpurun_out/nms_time.jsonl
for cap in 4 3 2; do
  VDET_NMS_PER_SM=$cap timeout 120 python tools/nms_time.py 1000 300 30 40 | tee -a gpurun_out/nms_time.jsonl
done
for cap in 4 3; do
  VDET_NMS_PER_SM=$cap timeout 120 python tools/nms_time.py 125 300 30 40 | tee -a gpurun_out/nms_time.jsonl
  VDET_NMS_PER_SM=$cap timeout 120 python tools/nms_time.py 2000 100 30 40 | tee -a gpurun_out/nms_time.jsonl
  VDET_NMS_PER_SM=$cap timeout 120 python tools/nms_time.py 1000 500 30 20 | tee -a gpurun_out/nms_time.jsonl
done
timeout 120 python tools/nms_time.py 500 2000 30 5 | tee -a gpurun_out/nms_time.jsonl
timeout 120 python tools/nms_time.py 300 1000 30 5 | tee -a gpurun_out/nms_time.jsonl
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], d["roofline"]["kernels_ms"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["mode"])
PY
