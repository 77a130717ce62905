set -u
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python tools/nms_shapes.py | tee gpurun_out/nms_shapes.jsonl
