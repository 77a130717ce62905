#!/bin/bash
# Round 2, call F (1 GPU): final evidence -- tests, smoke, bench (+ reference arm), kernel table, ncu launch list and captures.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "pytest -m gpu"
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
    print("ceiling", e["box_ceiling"]["upload_only_ms"], e["box_ceiling"]["stage_plus_upload_ms"], e["box_ceiling"]["e2e_vs_stage_plus_upload"], "registered", e["registered_inputs"]["ms_per_step"], "pinned", e["pinned_resubmit"]["ms_per_step"])
    print("parity", d["parity"])
    print("configs", json.dumps(d.get("configs"))[:2200])
    print("adapters", json.dumps(d.get("adapters"))[:1500])
except Exception as ex:
    print("bench unreadable", ex)
P
tail -n 3 gpurun_out/bench.err
el "reference arm"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
el "kernel bench"
timeout 240 python tools/kernel_bench.py > gpurun_out/kernels.txt 2>> gpurun_out/f.err; cut -c1-150 gpurun_out/kernels.txt
el "ncu launch list (bench, extras off)"
VDET_BENCH_EXTRAS=0 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
el "ncu full: step kernels"
VDET_BENCH_EXTRAS=0 timeout 220 ncu --set full --clock-control none --import-source on -k regex:'nms_frames_kernel|link_frames_kernel|iou_matrix_f32|compact_keep|keep_offsets' \
    -s 8 -c 6 -o gpurun_out/prof_r02 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"; tail -n 2 gpurun_out/ncu_full.log
el "sanitizer (memcheck + racecheck) on the NMS / compaction / completion tests"
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_nms.py tests/test_gpu_temporal_tubelet.py -m gpu -q -x -p no:cacheprovider \
    -k "frames_vs or tied or ragged or compact or completion or streams_new" > gpurun_out/sanitizer_memcheck.log 2>&1; tail -n 4 gpurun_out/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_nms.py -m gpu -q -x -p no:cacheprovider \
    -k "frames_vs or tied or ragged or compact" > gpurun_out/sanitizer_racecheck.log 2>&1; tail -n 4 gpurun_out/sanitizer_racecheck.log
el done
