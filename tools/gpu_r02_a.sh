#!/bin/bash
# Round 2, call A (1 GPU): measurements of the round-1 code that steer this round.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -30 >> gpurun_out/host.txt; numactl -H >> gpurun_out/host.txt 2>&1
el "small NMS: product vs pure_lds variant"
for i in 1 2; do
python tools/nms_time.py 1000 300 30 50 >> gpurun_out/nms_time_product.json 2>> gpurun_out/a.err
VDET_B200_LIB=vdetlib_b200/variants/libvdet_b200_pure_lds.so python tools/nms_time.py 1000 300 30 50 >> gpurun_out/nms_time_pure_lds.json 2>> gpurun_out/a.err
done
cat gpurun_out/nms_time_product.json gpurun_out/nms_time_pure_lds.json
el "variant parity"
VDET_B200_LIB=vdetlib_b200/variants/libvdet_b200_pure_lds.so timeout 200 python -m pytest tests/test_gpu_nms.py -m gpu -q -x -p no:cacheprovider -k "frames_vs or tied or golden or full_config2" > gpurun_out/pytest_pure_lds.log 2>&1; tail -n 2 gpurun_out/pytest_pure_lds.log
el "h2d probe N=1"
timeout 120 python tools/h2d_scale_probe.py > gpurun_out/h2d_probe_n1.json 2>> gpurun_out/a.err; cat gpurun_out/h2d_probe_n1.json
el "big kernel time"
python tools/run_big_nms.py 296 | tee gpurun_out/big_time.txt
el "ncu full: big kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nms_frames_big -s 2 -c 1 -o gpurun_out/prof_r02_big -f \
    python tools/run_big_nms.py 148 > gpurun_out/ncu_big.log 2>&1; tail -n 2 gpurun_out/ncu_big.log
el "ncu full: completion, iou f64, iou f32 (warm launches)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'completion_' -s 8 -c 2 -o gpurun_out/prof_r02_completion -f \
    python tools/kernel_bench.py --quick > gpurun_out/ncu_completion.log 2>&1; tail -n 2 gpurun_out/ncu_completion.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'iou_matrix_f' -s 6 -c 1 -o gpurun_out/prof_r02_iou32 -f \
    python tools/kernel_bench.py --quick > gpurun_out/ncu_iou32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'iou_matrix_f64' -s 3 -c 1 -o gpurun_out/prof_r02_iou64 -f \
    python tools/kernel_bench.py --quick > gpurun_out/ncu_iou64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'link_frames' -s 30 -c 1 -o gpurun_out/prof_r02_link5 -f \
    python tools/kernel_bench.py --quick > gpurun_out/ncu_link5.log 2>&1
el "kernel bench"
timeout 300 python tools/kernel_bench.py > gpurun_out/kernels.txt 2>> gpurun_out/a.err; tail -n 25 gpurun_out/kernels.txt
el "compute-sanitizer memcheck (small-frame + big-frame NMS parity tests)"
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_nms.py -m gpu -q -x -p no:cacheprovider \
    -k "frames_vs or tied or ragged" > gpurun_out/sanitizer_memcheck.log 2>&1; tail -n 6 gpurun_out/sanitizer_memcheck.log
el "compute-sanitizer racecheck"
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_nms.py -m gpu -q -x -p no:cacheprovider \
    -k "frames_vs or tied or ragged" > gpurun_out/sanitizer_racecheck.log 2>&1; tail -n 6 gpurun_out/sanitizer_racecheck.log
ls -la gpurun_out | head -40
el done
