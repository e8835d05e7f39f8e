#!/bin/bash
# Replicas: one native server + one load-generator instance per GPU (no collective anywhere).
#   gpurun --gpus N -- bash scripts/perf_native_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
OUT=gpurun_out/perf_native_multi_n$N.txt
: > $OUT
if which nvidia-cuda-mps-control > /dev/null 2>&1; then
  export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
  mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
  timeout 30 nvidia-cuda-mps-control -d
  sleep 2
  echo "## under CUDA MPS" >> $OUT
fi
URLS=""
PIDS=""
for i in $(seq 0 $((N-1))); do
  python -m client_b200.testing.native_server --port $((18100+i)) --device $i > gpurun_out/native_server_$i.log 2>&1 &
  PIDS="$PIDS $!"
  URLS="$URLS,127.0.0.1:$((18100+i))"
done
URLS=${URLS#,}
sleep 10
for g in 1 $N; do
  echo "## native engine, $g GPU instance(s), concurrency per instance" >> $OUT
  timeout 300 python -m client_b200.perf -m densenet_onnx -u $URLS --shared-memory cuda --engine native --gpus $g --concurrency-range 64:256:4x -p 1000 -r 5 --json >> $OUT 2>&1
done
kill $PIDS
wait $PIDS 2>/dev/null
if [ -n "$CUDA_MPS_PIPE_DIRECTORY" ]; then echo quit | timeout 30 nvidia-cuda-mps-control; fi
python - <<PY
import json
for line in open("$OUT"):
    if line.startswith("{"):
        r = json.loads(line)
        print("  gpus %s conc/instance %4d  %9.0f infer/s  p50 %7.1f us" % (r.get("gpus", 1), r["concurrency"], r["throughput"], r["p50_us"]))
    else:
        print(line.rstrip()[:200])
PY
