#!/bin/bash
# Loopback load runs on one GPU: mock server (own process) + the load generator in its modes.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python -m client_b200.testing.mock_server --http-port 18000 --grpc-port 18001 > gpurun_out/mock_server.log 2>&1 &
SRV=$!
sleep 6
{
echo "## C2 densenet cuda-shm, python engine"
timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18000 --shared-memory cuda --concurrency-range 1:64:4x -p 1000 -r 5 --json
echo "## C2 densenet cuda-shm, native engine"
timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18000 --shared-memory cuda --engine native --concurrency-range 1:64:4x -p 1000 -r 5 --json
echo "## C2 densenet cuda-shm, native engine, data filled once (perf_analyzer semantics)"
timeout 300 python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:18000 --shared-memory cuda --engine native --input-data-mode once --concurrency-range 64 -p 1000 -r 5 --json
echo "## C4 bert gRPC raw_input_contents (wire), python engine"
timeout 300 python -m client_b200.perf -m bert_large -u 127.0.0.1:18001 -i grpc --shared-memory none --concurrency-range 16:64:4x -p 1000 -r 5 --json
echo "## C5 llama decode stream, TTFT"
timeout 300 python -m client_b200.perf -m llama3_8b -u 127.0.0.1:18001 -i grpc --streaming --shape input_ids:1,4096 --concurrency-range 8 -p 1000 -r 5 --json
echo "## C1 simple, system shm, HTTP"
timeout 300 python -m client_b200.perf -m simple -u 127.0.0.1:18000 --shared-memory system --concurrency-range 4 -p 1000 -r 5 --json
echo "## reference-style CPU client loop (same server)"
timeout 120 python scripts/cpu_client_baseline.py -u 127.0.0.1:18000 --concurrency 8 --seconds 3
timeout 120 python scripts/cpu_client_baseline.py -u 127.0.0.1:18000 --concurrency 1 --seconds 3
} > gpurun_out/perf_loopback.txt 2>&1
kill $SRV
{
echo "## generator capacity: native engine against the canned-response server, C2 slots regenerated per request"
timeout 120 python - <<'PY'
import json
from client_b200.perf.loadgen import SlotSet, TensorSpec
from client_b200.perf.native import NativeLoadGenerator, StubServer
stub = StubServer()
for conc, regen in ((64, True), (64, False), (256, True)):
    ss = SlotSet([TensorSpec("data_0", "FP32", [3, 224, 224])], [TensorSpec("fc6_1", "FP32", [1000])], conc, "cuda", 0, "random", 1, name_prefix="cap%d%d" % (conc, regen))
    gen = NativeLoadGenerator(stub.url, "densenet_onnx", "", ss, conc, regenerate=regen, validate=True)
    gen.start(); gen.window(0.5); w = gen.window(2.0); gen.stop(); ss.close()
    w.update(concurrency=conc, regenerate=regen, input_gbps=w["throughput"] * 602112 / 1e9)
    print(json.dumps(w))
stub.stop()
PY
} >> gpurun_out/perf_loopback.txt 2>&1
cat gpurun_out/perf_loopback.txt | cut -c1-400
