"""Generator capacity over the gRPC transport against the canned-response gRPC stub (no model,
no second CUDA context): the C4 / C5 wire shapes and C2 over CUDA shared memory."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from client_b200.perf.loadgen import SlotSet, TensorSpec
from client_b200.perf.native import GrpcStubServer, NativeLoadGenerator, grpc_wire_prefixes

c2 = ([TensorSpec("data_0", "FP32", [3, 224, 224])], [TensorSpec("fc6_1", "FP32", [1000])])
c4 = ([TensorSpec("input_ids", "INT64", [1, 384]), TensorSpec("attention_mask", "INT64", [1, 384])], [TensorSpec("logits", "FP32", [1, 2])])
c5 = ([TensorSpec("input_ids", "INT32", [1, 4096])], [TensorSpec("logits", "FP32", [1, 16])])
levels = [int(x) for x in sys.argv[1:]] or [64, 256]
gstub = GrpcStubServer(b"\x0a\x01m")
for label, (ins, outs), shm in (("C4 bert gRPC raw_input_contents from pinned staging", c4, "none"), ("C5 llama prompt gRPC", c5, "none"),
                                ("C2 cuda-shm over gRPC", c2, "cuda")):
    for conc in levels:
        ss = SlotSet(ins, outs, conc, shm, 0, "random", 1, {"input_ids": (0, 30522), "attention_mask": (0, 2)}, name_prefix="gcap%s%d" % (shm, conc),
                     wire_prefixes=grpc_wire_prefixes(ins) if shm == "none" else None)
        gen = NativeLoadGenerator(gstub.url, "m", "", ss, conc, regenerate=True, validate=(shm == "cuda"), protocol="grpc")
        gen.start()
        gen.window(0.5)
        w = gen.window(2.0)
        gen.stop()
        ss.close()
        print(json.dumps({"case": label, "concurrency": conc, "infer_per_s": round(w["throughput"]), "p50_us": w["p50_us"], "failed": w["failed"],
                          "slots_per_pass": round(w["device_slots"] / max(1, w["device_batches"]), 1),
                          "input_gbps": round(w["throughput"] * ss.in_bytes / 1e9, 2)}), flush=True)
gstub.stop()
# C5 as BASELINE quotes it: a ModelStreamInfer stream per connection, the decoupled model answers with 16 tokens
from client_b200.perf.native import stream_token_responses
resp, fin = stream_token_responses()
sstub = GrpcStubServer(resp, final_response=fin, responses_per_request=16)
ins, outs = c5
for conc in levels:
    ss = SlotSet(ins, [], conc, "none", 0, "random", 1, {"input_ids": (0, 128256)}, name_prefix="gstream%d" % conc, wire_prefixes=grpc_wire_prefixes(ins))
    gen = NativeLoadGenerator(sstub.url, "llama3_8b", "", ss, conc, regenerate=True, validate=False, protocol="grpc-stream")
    gen.start()
    gen.window(0.5)
    w = gen.window(2.0)
    gen.stop()
    ss.close()
    print(json.dumps({"case": "C5 llama prompt on a ModelStreamInfer stream, 16 token responses per request", "concurrency": conc,
                      "infer_per_s": round(w["throughput"]), "tokens_per_s": round(w["responses_per_s"]), "ttft_p50_us": w["ttft_p50_us"],
                      "ttft_p99_us": w["ttft_p99_us"], "p50_us": w["p50_us"], "failed": w["failed"],
                      "slots_per_pass": round(w["device_slots"] / max(1, w["device_batches"]), 1)}), flush=True)
sstub.stop()
