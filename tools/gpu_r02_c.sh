#!/bin/bash
# Round 2, call C (1 GPU): new big-kernel walk, registered inputs, host breakdown, register probe, fresh ncu of the small kernel.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "big kernel time"
timeout 90 python tools/run_big_nms.py 296 | tee gpurun_out/big_time.txt; echo "big rc=$?"
el "pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 150 --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
el "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
el "bench"
timeout 420 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.4g (%.4f ms) nms %.4f link %.4f iou_frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"]["nms_frames_kernel"], d["roofline"]["kernels_ms"]["link_frames_kernel"], d["iou_matrix_roofline"]["frac"]))
print("e2e %.4g (%.4f ms) wall %.4f host %s registered %s pinned %s" % (e["value"], e["ms_per_step"], e["host_wall_ms_per_step"], e["host_ms_per_step"], e["registered_inputs"], e["pinned_resubmit"]["ms_per_step"]))
print("parity", d["parity"])
print("configs", json.dumps(d.get("configs"))[:1500])
print("adapters", json.dumps(d.get("adapters"))[:1500])
P
tail -n 5 gpurun_out/bench.err
el "h2d probe with register legs"
timeout 120 python tools/h2d_scale_probe.py --reps 30 > gpurun_out/h2d_probe_n1.json 2>> gpurun_out/c.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/h2d_probe_n1.json").read())
print({k: v["ms_per_round_max"] for k, v in d["legs"].items()})
P
el "ncu launch list + full capture of the step's kernels (bench, extras off)"
VDET_BENCH_EXTRAS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
VDET_BENCH_EXTRAS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'nms_frames_kernel|link_frames_kernel|iou_matrix_f32|compact_keep|keep_offsets' \
    -s 8 -c 6 -o gpurun_out/prof_r02 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"; tail -n 2 gpurun_out/ncu_full.log
el "ncu full: big kernel + completion"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nms_frames_big -s 2 -c 1 -o gpurun_out/prof_r02_big3 -f \
    python tools/run_big_nms.py 148 > gpurun_out/ncu_big3.log 2>&1; tail -n 1 gpurun_out/ncu_big3.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'completion_' -s 8 -c 2 -o gpurun_out/prof_r02_completion2 -f \
    python tools/kernel_bench.py --quick > gpurun_out/ncu_completion2.log 2>&1; tail -n 1 gpurun_out/ncu_completion2.log
el done
