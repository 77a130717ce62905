#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): host<->device ceiling of the box at 1/2/4/N ranks, multi-rank parity
# tests, bench.py at every N (both arms at N=1 only: the reference arm does not change with N).
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_r02_multi.sh 8 [noprobe]'
set -u
NG=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi topo -m > gpurun_out/topo_n$NG.txt 2>&1; nproc >> gpurun_out/topo_n$NG.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/topo_n$NG.txt
PORT=29610
run_n() {  # $1 ranks, rest: script + args
  local n=$1; shift
  PORT=$((PORT + 1))
  if [ "$n" = "1" ]; then python "$@"; else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT "$@"; fi
}
el "h2d probe"
[ "${2:-}" = "noprobe" ] && SKIP_PROBE=1 || SKIP_PROBE=0
for n in 1 2 4 8; do
  [ $SKIP_PROBE = 1 ] && continue
  [ $n -le $NG ] || continue
  timeout 120 bash -c "$(declare -f run_n); PORT=$((29700 + n)); run_n $n tools/h2d_scale_probe.py --reps 30" >> gpurun_out/h2d_probe_scale.json 2>> gpurun_out/multi.err
  if [ $n -ge 4 ]; then
    timeout 120 bash -c "$(declare -f run_n); PORT=$((29750 + n)); run_n $n tools/h2d_scale_probe.py --reps 30 --bind" >> gpurun_out/h2d_probe_scale.json 2>> gpurun_out/multi.err
  fi
done
[ $SKIP_PROBE = 1 ] || python - <<'P'
import json
for line in open("gpurun_out/h2d_probe_scale.json"):
    try:
        d = json.loads(line)
    except Exception:
        continue
    print("ranks", d["ranks"], "bind", d["bind"], "thr", d["stage_threads"],
          {k: (v["ms_per_round_max"], v["aggregate_up_GBs"]) for k, v in d["legs"].items()})
P
el "multi-rank parity tests"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 180 -p no:cacheprovider > gpurun_out/pytest_multi_n$NG.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_multi_n$NG.log
el "bench"
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  el "bench N=$n"
  VDET_BENCH_EXTRAS=0 timeout 300 bash -c "$(declare -f run_n); PORT=$((29800 + n)); run_n $n bench.py --gpus $n --steps 50 --warmup 5" > gpurun_out/bench_n$n.json 2>> gpurun_out/multi.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    c = e.get("box_ceiling", {})
    print("N=$n value %.4g (%.4f ms)  e2e %.4g (%.3f ms)  ceiling stage+up %.3f up %.3f  registered %.3f  pinned %.3f ms  parity %s %s" % (
        d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], c.get("stage_plus_upload_ms", 0), c.get("upload_only_ms", 0),
        e["registered_inputs"]["ms_per_step"], e["pinned_resubmit"]["ms_per_step"],
        all(v for k, v in d["parity"].items() if isinstance(v, bool)), d.get("parity_multi")))
except Exception as ex:
    print("bench N=$n unreadable:", ex)
P
done
tail -n 20 gpurun_out/multi.err
el done
