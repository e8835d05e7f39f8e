#!/bin/bash
# C2 (densenet_onnx over CUDA shared memory) against the native server with and without look-ahead,
# time-sliced and under MPS.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
round() {
  python -m client_b200.testing.native_server --port $1 > gpurun_out/native_server_la_$2.log 2>&1 &
  SRV=$!
  sleep 8
  for la in 1 4 8; do
    echo "## C2 densenet cuda-shm, native engine, --lookahead $la"
    timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:$1 --shared-memory cuda --engine native --lookahead $la --concurrency-range 1:256:4x -p 800 -r 4 --json
  done
  kill $SRV
  wait $SRV 2>/dev/null
}
{
echo "## === time-sliced contexts (no MPS) ==="
round 18300 plain
if which nvidia-cuda-mps-control > /dev/null 2>&1; then
  export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
  mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
  timeout 30 nvidia-cuda-mps-control -d
  sleep 2
  echo "## === under CUDA MPS ==="
  round 18310 mps
  echo quit | timeout 30 nvidia-cuda-mps-control
fi
} > gpurun_out/lookahead.txt 2>&1
python - <<PY
import json
for line in open("gpurun_out/lookahead.txt"):
    if line.startswith("##"): print(line.strip()); continue
    if line.startswith("{"):
        d = json.loads(line)
        print("   conc %3d: %8.0f /s p50 %7.1f us p99 %7.1f slots/pass %.1f failed %d nonfinite %d" % (d["concurrency"], d["throughput"], d["p50_us"], d["p99_us"], d["device_slots"] / max(1, d["device_batches"]), d["failed"], d["nonfinite"]))
    elif line.strip(): print(line.strip()[:200])
PY
