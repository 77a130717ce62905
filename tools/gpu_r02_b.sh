#!/bin/bash
# Round 2, call B (1 GPU): parity tests + smoke + bench of the rebuilt e2e path, the rewritten big-frame NMS kernel.
# Every step has its own short timeout: a hanging kernel must not eat the call.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "big kernel time (first: a hang shows here within a minute)"
timeout 90 python tools/run_big_nms.py 296 | tee gpurun_out/big_time.txt; echo "big rc=$?"
el "pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 150 --tb=short -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -n 40 gpurun_out/pytest_gpu.log
el "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
el "bench"
timeout 420 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -n 8 gpurun_out/bench.err
el "small NMS: product (packed tile) vs scalar tile vs pure_lds"
for i in 1 2; do
timeout 60 python tools/nms_time.py 1000 300 30 50 >> gpurun_out/nms_time_product.json 2>> gpurun_out/b.err
VDET_B200_LIB=vdetlib_b200/variants/libvdet_b200_scalar_tile.so timeout 60 python tools/nms_time.py 1000 300 30 50 >> gpurun_out/nms_time_scalar_tile.json 2>> gpurun_out/b.err
VDET_B200_LIB=vdetlib_b200/variants/libvdet_b200_pure_lds.so timeout 60 python tools/nms_time.py 1000 300 30 50 >> gpurun_out/nms_time_pure_lds.json 2>> gpurun_out/b.err
done
tail -n 2 gpurun_out/nms_time_product.json gpurun_out/nms_time_scalar_tile.json gpurun_out/nms_time_pure_lds.json
el "h2d probe: sse vs avx512, thread sweep"
for th in 4 8 16; do
  VDET_HOST_COPY=avx512 timeout 90 python tools/h2d_scale_probe.py --threads $th --reps 30 >> gpurun_out/h2d_probe_avx512.json 2>> gpurun_out/b.err
done
VDET_HOST_COPY=sse timeout 90 python tools/h2d_scale_probe.py --threads 8 --reps 30 >> gpurun_out/h2d_probe_sse.json 2>> gpurun_out/b.err
python - <<'P'
import json
for f in ("gpurun_out/h2d_probe_avx512.json", "gpurun_out/h2d_probe_sse.json"):
    try:
        for line in open(f):
            d = json.loads(line)
            print(f.split("_")[-1], d["stage_threads"], {k: v["ms_per_round_max"] for k, v in d["legs"].items()})
    except Exception as e:
        print(f, e)
P
el "kernel bench"
timeout 240 python tools/kernel_bench.py > gpurun_out/kernels.txt 2>> gpurun_out/b.err; cat gpurun_out/kernels.txt | cut -c1-150
el "ncu full: big kernel (rewritten)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:nms_frames_big -s 2 -c 1 -o gpurun_out/prof_r02_big2 -f \
    python tools/run_big_nms.py 148 > gpurun_out/ncu_big2.log 2>&1; tail -n 2 gpurun_out/ncu_big2.log
el done
