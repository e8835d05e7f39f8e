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
{
echo "## generator capacity against the canned-response server (no model, no second CUDA context)"
timeout 200 python - <<'PY'
import json
from client_b200.perf.loadgen import SlotSet, TensorSpec
from client_b200.perf.native import NativeLoadGenerator, StubServer
stub = StubServer()
c2 = ([TensorSpec("data_0", "FP32", [3, 224, 224])], [TensorSpec("fc6_1", "FP32", [1000])])
c4 = ([TensorSpec("input_ids", "INT64", [1, 384]), TensorSpec("attention_mask", "INT64", [1, 384])], [TensorSpec("logits", "FP32", [1, 2])])
c5 = ([TensorSpec("input_ids", "INT32", [1, 4096])], [TensorSpec("logits", "FP32", [1, 16])])
for label, (ins, outs), shm, conc in (("C2 cuda-shm", c2, "cuda", 64), ("C2 cuda-shm", c2, "cuda", 256),
                                      ("C4 bert wire (HTTP binary body from pinned staging)", c4, "none", 256),
                                      ("C5 llama prompt wire", c5, "none", 256),
                                      ("C2 wire (602 KB bodies)", c2, "none", 64)):
    ss = SlotSet(ins, outs, conc, shm, 0, "random", 1, {"input_ids": (0, 30522), "attention_mask": (0, 2)}, name_prefix="cap%s%d" % (shm, conc))
    gen = NativeLoadGenerator(stub.url, "m", "", ss, conc, regenerate=True, validate=(shm == "cuda"))
    gen.start(); gen.window(0.5); w = gen.window(2.0); gen.stop(); ss.close()
    print(json.dumps({"case": label, "concurrency": conc, "infer_per_s": round(w["throughput"]), "p50_us": w["p50_us"], "failed": w["failed"],
                      "slots_per_pass": round(w["device_slots"] / max(1, w["device_batches"]), 1),
                      "input_gbps": round(w["throughput"] * ss.in_bytes / 1e9, 2)}))
stub.stop()
# the same generator over its gRPC transport (cleartext HTTP/2): C4 / C5 are quoted on gRPC
from client_b200.perf.native import GrpcStubServer, grpc_wire_prefixes
gstub = GrpcStubServer(b"\x0a\x01m")
for label, (ins, outs), shm, conc in (("C4 bert gRPC raw_input_contents from pinned staging", c4, "none", 256),
                                      ("C5 llama prompt gRPC", c5, "none", 256),
                                      ("C2 cuda-shm over gRPC", c2, "cuda", 256)):
    ss = SlotSet(ins, outs, conc, shm, 0, "random", 1, {"input_ids": (0, 30522), "attention_mask": (0, 2)}, name_prefix="gcap%s%d" % (shm, conc),
                 wire_prefixes=grpc_wire_prefixes(ins) if shm == "none" else None)
    gen = NativeLoadGenerator(gstub.url, "m", "", ss, conc, regenerate=True, validate=(shm == "cuda"), protocol="grpc")
    gen.start(); gen.window(0.5); w = gen.window(2.0); gen.stop(); ss.close()
    print(json.dumps({"case": label, "concurrency": conc, "infer_per_s": round(w["throughput"]), "p50_us": w["p50_us"], "failed": w["failed"],
                      "slots_per_pass": round(w["device_slots"] / max(1, w["device_batches"]), 1),
                      "input_gbps": round(w["throughput"] * ss.in_bytes / 1e9, 2)}))
gstub.stop()
PY
} >> gpurun_out/perf_native.txt 2>&1
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
