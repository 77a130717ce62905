#!/bin/bash
# Round 2, call I (1 GPU): the round's last state -- all GPU tests, smoke, bench, reference arm, kernel table, the
# frame-count sweep of the two NMS CTA shapes, one ncu capture of the NMS kernel, launch list, racecheck of the NMS tests.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "pytest -m gpu"
timeout 200 python -m pytest tests -m gpu -q -x --timeout 100 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu.log
el "smoke"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
el "bench"
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value %.4g (%.4f ms) nms %.4f link %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernels_ms"]["nms_frames_kernel"], d["roofline"]["kernels_ms"]["link_frames_kernel"]))
    print("e2e %.4g (%.4f ms) registered %.4f pinned %.4f" % (e["value"], e["ms_per_step"], e["registered_inputs"]["ms_per_step"], e["pinned_resubmit"]["ms_per_step"]))
    print("parity", d["parity"])
    c = d["configs"]
    print("C3 link ms", c["config3_link"]["ms"], "C5", c["config5_video"]["kernels_ms"])
except Exception as ex:
    print("bench unreadable", ex)
P
tail -n 2 gpurun_out/bench.err
el "reference arm"
timeout 100 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_ref.json
el "kernel bench"
timeout 120 python tools/kernel_bench.py > gpurun_out/kernels.txt 2>> gpurun_out/i.err; head -n 3 gpurun_out/kernels.txt | cut -c1-150
el "frame-count sweep of the CTA shapes"
timeout 120 python tools/nms_shapes.py 125,300,30 300,300,30 450,300,30 600,300,30 800,300,30 1000,300,30 1184,300,30 1500,300,30 2000,300,30 3000,300,30 1000,300,20 > gpurun_out/nms_shapes_T.jsonl 2>> gpurun_out/i.err; cat gpurun_out/nms_shapes_T.jsonl
el "ncu full: NMS kernel"
VDET_BENCH_EXTRAS=0 timeout 120 ncu --set full --clock-control none --import-source on -k regex:'nms_frames_kernel' \
    -s 8 -c 1 -o gpurun_out/prof_r02_nms2 -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
el "ncu launch list"
VDET_BENCH_EXTRAS=0 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches2.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
el "racecheck"
timeout 100 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_nms.py -m gpu -q -x -p no:cacheprovider \
    -k "frames_vs and (17-300 or 9-33 or 4-150) or integer and 6-300-5-0.5 or ragged_frames" > gpurun_out/sanitizer_racecheck2.log 2>&1; tail -n 3 gpurun_out/sanitizer_racecheck2.log
el done
