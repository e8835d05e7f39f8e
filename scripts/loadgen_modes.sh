#!/bin/bash
# Experiment: accumulation windows of the native load generator and the native server.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
: > gpurun_out/loadgen_modes.txt
for sdelay in 0 100; do
  TB200_SERVER_QUEUE_DELAY_US=$sdelay python -m client_b200.testing.native_server --port 18100 > gpurun_out/native_server.log 2>&1 &
  SRV=$!
  sleep 7
  for win in 0 50 150; do
    for mode in 0 1; do
      echo "## server_delay=$sdelay client_window=$win mode=$mode" >> gpurun_out/loadgen_modes.txt
      TB200_LOADGEN_WINDOW_US=$win TB200_LOADGEN_DEVICE_MODE=$mode timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18100 --shared-memory cuda --engine native --concurrency-range 1:256:4x -p 500 -r 3 --json >> gpurun_out/loadgen_modes.txt 2>&1
    done
  done
  kill $SRV
  wait $SRV 2>/dev/null
done
python - <<'PY'
import json
for line in open("gpurun_out/loadgen_modes.txt"):
    if line.startswith("##"):
        print(line.strip())
    elif line.startswith("{"):
        r = json.loads(line)
        print("  conc %4d  %9.0f infer/s  p50 %7.1f us  p99 %8.1f us  slots/pass %.1f" % (r["concurrency"], r["throughput"], r["p50_us"], r["p99_us"], r["device_slots"] / max(1, r["device_batches"])))
PY
