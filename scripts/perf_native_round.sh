#!/bin/bash
# Sustained inferences/sec against the NATIVE local CUDA-shared-memory server (own process):
# device-side load generator vs. the reference-style CPU client loop, same server, same model.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python -m client_b200.testing.native_server --port 18100 > gpurun_out/native_server.log 2>&1 &
SRV=$!
sleep 8
{
echo "## C2 densenet cuda-shm vs native server: native engine, inputs regenerated + outputs validated per request"
timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18100 --shared-memory cuda --engine native --concurrency-range 1:256:4x -p 1000 -r 5 --json
echo "## same with --device-window-us 150 (full passes; steadier without MPS)"
timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18100 --shared-memory cuda --engine native --device-window-us 150 --concurrency-range 16:256:4x -p 1000 -r 5 --json
echo "## same, data filled once (perf_analyzer semantics)"
timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18100 --shared-memory cuda --engine native --input-data-mode once --concurrency-range 64 -p 1000 -r 5 --json
echo "## python engine (client_b200.http + device fill/check)"
timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18100 --shared-memory cuda --concurrency-range 1:16:4x -p 1000 -r 5 --json
echo "## reference-style CPU client loop: 1 thread, 8 threads, 8/32 processes"
timeout 120 python scripts/cpu_client_baseline.py -u 127.0.0.1:18100 --concurrency 1 --seconds 3
timeout 120 python scripts/cpu_client_baseline.py -u 127.0.0.1:18100 --concurrency 8 --seconds 3
timeout 200 python scripts/cpu_client_baseline.py -u 127.0.0.1:18100 --concurrency 1 --processes 8 --seconds 4
timeout 300 python scripts/cpu_client_baseline.py -u 127.0.0.1:18100 --concurrency 1 --processes 32 --seconds 4
} > gpurun_out/perf_native.txt 2>&1
kill $SRV
wait $SRV 2>/dev/null
cat gpurun_out/native_server.log >> gpurun_out/perf_native.txt
# ---- the same comparison with client and server sharing the GPU through CUDA MPS
if which nvidia-cuda-mps-control > /dev/null 2>&1; then
  export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
  mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
  timeout 30 nvidia-cuda-mps-control -d
  sleep 2
  python -m client_b200.testing.native_server --port 18101 > gpurun_out/native_server_mps.log 2>&1 &
  SRV=$!
  sleep 8
  {
  echo "## === under CUDA MPS (client and server contexts run concurrently) ==="
  echo "## native engine, inputs regenerated + outputs validated per request"
  timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18101 --shared-memory cuda --engine native --concurrency-range 1:256:4x -p 1000 -r 5 --json
  echo "## python engine"
  timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18101 --shared-memory cuda --concurrency-range 1:16:4x -p 1000 -r 5 --json
  echo "## reference-style CPU client loop: 1 thread, 8 / 32 processes"
  timeout 120 python scripts/cpu_client_baseline.py -u 127.0.0.1:18101 --concurrency 1 --seconds 3
  timeout 200 python scripts/cpu_client_baseline.py -u 127.0.0.1:18101 --concurrency 1 --processes 8 --seconds 4
  timeout 300 python scripts/cpu_client_baseline.py -u 127.0.0.1:18101 --concurrency 1 --processes 32 --seconds 4
  } >> gpurun_out/perf_native.txt 2>&1
  kill $SRV
  wait $SRV 2>/dev/null
  echo quit | timeout 30 nvidia-cuda-mps-control
  unset CUDA_MPS_PIPE_DIRECTORY CUDA_MPS_LOG_DIRECTORY
fi
cut -c1-420 gpurun_out/perf_native.txt
