#!/bin/bash
# Loopback run with client and server sharing the GPU through CUDA MPS (if the image has it):
# without MPS the two contexts are time-sliced and every hand-over costs ~70-100 us.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
OUT=gpurun_out/mps_round.txt
: > $OUT
MPS=$(which nvidia-cuda-mps-control 2>/dev/null)
echo "nvidia-cuda-mps-control: ${MPS:-not found}" >> $OUT
if [ -z "$MPS" ]; then cat $OUT; exit 0; fi
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
timeout 30 nvidia-cuda-mps-control -d >> $OUT 2>&1
sleep 2
python -m client_b200.testing.native_server --port 18100 > gpurun_out/native_server.log 2>&1 &
SRV=$!
sleep 8
for win in 0 150; do
  for mode in 0 1; do
    echo "## MPS client_window=$win mode=$mode" >> $OUT
    TB200_LOADGEN_WINDOW_US=$win TB200_LOADGEN_DEVICE_MODE=$mode timeout 200 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18100 --shared-memory cuda --engine native --concurrency-range 1:256:4x -p 500 -r 3 --json >> $OUT 2>&1
  done
done
echo "## MPS reference-style CPU loop, 8 processes" >> $OUT
timeout 200 python scripts/cpu_client_baseline.py -u 127.0.0.1:18100 --concurrency 1 --processes 8 --seconds 3 >> $OUT 2>&1
timeout 60 python scripts/cpu_client_baseline.py -u 127.0.0.1:18100 --concurrency 1 --seconds 3 >> $OUT 2>&1
kill $SRV
wait $SRV 2>/dev/null
echo quit | timeout 30 nvidia-cuda-mps-control >> $OUT 2>&1
tail -5 /tmp/mps_log/control.log >> $OUT 2>/dev/null
python - <<'PY'
import json
for line in open("gpurun_out/mps_round.txt"):
    if line.startswith("{") and "device_slots" in line:
        r = json.loads(line)
        print("  conc %4d  %9.0f infer/s  p50 %7.1f us  p99 %8.1f us  slots/pass %.1f" % (r["concurrency"], r["throughput"], r["p50_us"], r["p99_us"], r["device_slots"] / max(1, r["device_batches"])))
    else:
        print(line.rstrip()[:300])
PY
