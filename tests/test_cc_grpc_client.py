"""C++ gRPC front end (client_b200/cpp/tb200_grpc_client.h): message classes
generated without protoc, own HTTP/2 framing.  The C++ codec is checked against libprotobuf
(through the Python runtime): byte-identical serialisation of the requests the client forms,
byte-identical re-serialisation and text format of randomly filled messages; the client itself
runs against the grpcio mock server (tests/cpp/test_cc_grpc_client.cc)."""

import os
import random
import subprocess
import sys

import numpy as np
import pytest

from test_loopback import start_server

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "client_b200", "lib")


@pytest.fixture(scope="module")
def binary():
    from client_b200.build import build_cpp_client, build_native

    build_native()
    build_cpp_client()
    out = os.path.join(ROOT, "build", "test_cc_grpc_client")
    cpp = os.path.join(ROOT, "client_b200", "cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I" + os.path.join(cpp, "compat"), "-I" + cpp,
                    os.path.join(ROOT, "tests", "cpp", "test_cc_grpc_client.cc"), "-o", out,
                    "-L" + LIBDIR, "-ltb200client", "-ltb200", "-Wl,-rpath," + LIBDIR, "-lpthread"], check=True)
    return out


def test_generated_header_is_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_pb_cpp.py"), "--check"], cwd=ROOT)
    assert r.returncode == 0, "client_b200/cpp/grpc_service.pb.h is stale: run python scripts/gen_pb_cpp.py"


def test_known_answers_offline(binary):
    r = subprocess.run([binary], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASS (offline)" in r.stdout, r.stdout + r.stderr


def test_requests_match_libprotobuf(binary):
    """The ModelInferRequest bytes the C++ client sends == the bytes libprotobuf produces for the
    message the reference C++ client builds (grpc_client.cc:1419-1580: parameters incl. the always
    present triton_enable_empty_final_response, shared-memory parameters, raw_input_contents)."""
    from client_b200.grpc import service_pb2

    r = subprocess.run([binary, "--requests"], capture_output=True, text=True, timeout=60)
    got = dict(line.split() for line in r.stdout.strip().splitlines())
    x = np.arange(16, dtype=np.int32).tobytes()
    y = np.full(16, -1, np.int32).tobytes()

    def base(model, ins, outs):
        m = service_pb2.ModelInferRequest(model_name=model)
        for name in ins:
            t = m.inputs.add()
            t.name, t.datatype = name, "INT32"
            t.shape.extend([1, 16])
        for name in outs:
            m.outputs.add().name = name
        return m

    plain = base("simple", ["INPUT0", "INPUT1"], ["OUTPUT0", "OUTPUT1"])
    plain.parameters["triton_enable_empty_final_response"].bool_param = False
    plain.outputs[1].parameters["classification"].int64_param = 3
    plain.raw_input_contents.extend([x, y])
    assert got["plain"] == plain.SerializeToString(deterministic=True).hex()

    opt = base("simple", ["INPUT0", "INPUT1"], ["OUTPUT0"])
    opt.model_version, opt.id = "2", "req-7"
    opt.parameters["triton_enable_empty_final_response"].bool_param = True
    opt.parameters["sequence_id"].int64_param = 1007
    opt.parameters["sequence_start"].bool_param = True
    opt.parameters["sequence_end"].bool_param = False
    opt.parameters["priority"].uint64_param = 3
    opt.parameters["timeout"].int64_param = 5000
    opt.parameters["my_key"].string_param = "v"
    opt.parameters["count"].int64_param = -12
    opt.parameters["flag"].bool_param = True
    opt.raw_input_contents.extend([x, y])
    assert got["options"] == opt.SerializeToString(deterministic=True).hex()

    shm = base("simple", ["INPUT0", "INPUT1"], ["OUTPUT0", "OUTPUT1"])
    shm.parameters["triton_enable_empty_final_response"].bool_param = False
    shm.parameters["sequence_id"].string_param = "seq-a"
    shm.parameters["sequence_start"].bool_param = False
    shm.parameters["sequence_end"].bool_param = True
    for t, region, off in ((shm.inputs[0], "input_data", 0), (shm.inputs[1], "input_data", 64),
                           (shm.outputs[0], "output_data", 0), (shm.outputs[1], "output_data", 64)):
        t.parameters["shared_memory_region"].string_param = region
        t.parameters["shared_memory_byte_size"].int64_param = 64
        if off:
            t.parameters["shared_memory_offset"].int64_param = off
    shm.outputs[1].parameters["classification"].int64_param = 3
    assert got["shm"] == shm.SerializeToString(deterministic=True).hex()


def _random_messages(rng):
    from client_b200.grpc import model_config_pb2, service_pb2

    def text(n=8):
        return "".join(rng.choice("abcXYZ019_-/ \"'\\\n\té") for _ in range(rng.randrange(n)))

    def blob(n=24):
        return bytes(rng.randrange(256) for _ in range(rng.randrange(n)))

    def param(p):
        k = rng.randrange(5)
        if k == 0:
            p.bool_param = rng.random() < 0.5
        elif k == 1:
            p.int64_param = rng.choice([0, -1, 1, 2**62, -2**63, rng.randrange(-10**6, 10**6)])
        elif k == 2:
            p.string_param = text()
        elif k == 3:
            p.double_param = rng.choice([0.0, -0.0, 1.5, 1e300, 0.1, rng.random()])
        else:
            p.uint64_param = rng.choice([0, 2**64 - 1, rng.randrange(10**9)])

    out = []
    for _ in range(40):
        m = service_pb2.ModelInferRequest(model_name=text(), model_version=text(3), id=text(4))
        for _ in range(rng.randrange(4)):
            param(m.parameters[text() or "k"])
        for _ in range(rng.randrange(3)):
            t = m.inputs.add()
            t.name, t.datatype = text(), rng.choice(["INT32", "FP32", "BYTES", ""])
            t.shape.extend(rng.choice([-1, 0, 1, 224, 2**40]) for _ in range(rng.randrange(4)))
            for _ in range(rng.randrange(3)):
                param(t.parameters[text() or "p"])
            if rng.random() < 0.4:
                c = t.contents
                c.bool_contents.extend(rng.random() < 0.5 for _ in range(rng.randrange(4)))
                c.int_contents.extend(rng.choice([-2**31, 2**31 - 1, 0, -1, 7]) for _ in range(rng.randrange(4)))
                c.int64_contents.extend(rng.choice([-2**63, 2**63 - 1, 0, -1]) for _ in range(rng.randrange(4)))
                c.uint_contents.extend(rng.choice([0, 2**32 - 1, 5]) for _ in range(rng.randrange(4)))
                c.uint64_contents.extend(rng.choice([0, 2**64 - 1, 5]) for _ in range(rng.randrange(4)))
                c.fp32_contents.extend(rng.choice([0.0, -0.0, 1.5, 3.4e38, 0.1, float("inf")]) for _ in range(rng.randrange(4)))
                c.fp64_contents.extend(rng.choice([0.0, 1e-300, 0.1, 2.5, -7.25e10]) for _ in range(rng.randrange(4)))
                c.bytes_contents.extend(blob() for _ in range(rng.randrange(3)))
        for _ in range(rng.randrange(3)):
            o = m.outputs.add()
            o.name = text()
            for _ in range(rng.randrange(2)):
                param(o.parameters[text() or "q"])
        m.raw_input_contents.extend(blob(64) for _ in range(rng.randrange(3)))
        out.append(("ModelInferRequest", m))
    for _ in range(20):
        m = service_pb2.ModelInferResponse(model_name=text(), id=text())
        if rng.random() < 0.5:
            m.parameters["triton_final_response"].bool_param = rng.random() < 0.5
        for _ in range(rng.randrange(3)):
            o = m.outputs.add()
            o.name, o.datatype = text(), "FP32"
            o.shape.extend([1, rng.randrange(1000)])
        m.raw_output_contents.extend(blob(40) for _ in range(len(m.outputs)))
        out.append(("ModelInferResponse", m))
        s = service_pb2.ModelStreamInferResponse(error_message=text() if rng.random() < 0.3 else "")
        if rng.random() < 0.8:
            s.infer_response.CopyFrom(m)
        out.append(("ModelStreamInferResponse", s))
    for _ in range(10):
        c = service_pb2.ModelConfigResponse()
        c.config.name, c.config.platform, c.config.max_batch_size = text(), text(), rng.choice([0, 8, -1])
        for _ in range(rng.randrange(3)):
            i = c.config.input.add()
            i.name, i.data_type = text(), rng.randrange(15)
            i.format = rng.randrange(3)
            i.dims.extend([3, 224, -1][: rng.randrange(4)])
            i.optional = rng.random() < 0.5
        for _ in range(rng.randrange(3)):
            o = c.config.output.add()
            o.name, o.data_type, o.label_filename = text(), rng.randrange(15), text()
            o.dims.extend([1000])
        if rng.random() < 0.5:
            c.config.model_transaction_policy.decoupled = rng.random() < 0.5
        # the sections of model_config.proto beyond inputs / outputs: nested enums, oneofs
        # with message members, map<uint64, message>, map<string, string>, map<string, message>
        for _ in range(rng.randrange(3)):
            g = c.config.instance_group.add()
            g.name, g.kind, g.count = text(), rng.randrange(5), rng.randrange(4)
            g.gpus.extend(rng.randrange(8) for _ in range(rng.randrange(3)))
            if rng.random() < 0.4:
                r = g.rate_limiter.resources.add()
                r.name, r.count = text(), rng.randrange(9)
                setattr(r, "global", rng.random() < 0.5)
            for _ in range(rng.randrange(2)):
                sd = g.secondary_devices.add()
                sd.kind, sd.device_id = 0, rng.randrange(4)
        pick = rng.randrange(4)
        if pick == 0:
            c.config.dynamic_batching.preferred_batch_size.extend([4, 8][: rng.randrange(3)])
            c.config.dynamic_batching.max_queue_delay_microseconds = rng.randrange(10**6)
            for _ in range(rng.randrange(3)):
                q = c.config.dynamic_batching.priority_queue_policy[rng.randrange(1, 6)]
                q.timeout_action, q.max_queue_size, q.allow_timeout_override = rng.randrange(2), rng.randrange(100), rng.random() < 0.5
            if rng.random() < 0.5:
                c.config.dynamic_batching.default_queue_policy.default_timeout_microseconds = rng.randrange(10**9)
        elif pick == 1:
            sb = c.config.sequence_batching
            sb.max_sequence_idle_microseconds = rng.randrange(10**9)
            if rng.random() < 0.5:
                sb.oldest.max_candidate_sequences = rng.randrange(64)
                sb.oldest.preferred_batch_size.extend([2, 4])
            else:
                sb.direct.max_queue_delay_microseconds = rng.randrange(1000)
            ci = sb.control_input.add()
            ci.name = text()
            k = ci.control.add()
            k.kind = rng.randrange(4)
            k.int32_false_true.extend([0, 1])
            st8 = sb.state.add()
            st8.input_name, st8.output_name, st8.data_type = text(), text(), rng.randrange(15)
            st8.dims.extend([-1])
            if rng.random() < 0.5:
                st8.initial_state.add(name=text(), data_type=11, dims=[1], zero_data=True)
            else:
                st8.initial_state.add(name=text(), data_type=11, dims=[1], data_file=text())
        elif pick == 2:
            step = c.config.ensemble_scheduling.step.add()
            step.model_name, step.model_version = text(), rng.choice([-1, 1, 3])
            for _ in range(rng.randrange(3)):
                step.input_map[text() or "in"] = text()
                step.output_map[text() or "out"] = text()
        for _ in range(rng.randrange(3)):
            c.config.parameters[text() or "p"].string_value = text()
        if rng.random() < 0.5:
            c.config.version_policy.specific.versions.extend([1, 3, rng.randrange(10)])
        elif rng.random() < 0.5:
            c.config.version_policy.latest.num_versions = rng.randrange(4)
        if rng.random() < 0.4:
            c.config.optimization.cuda.graphs = True
            gs = c.config.optimization.cuda.graph_spec.add()
            gs.batch_size = rng.randrange(8)
            gs.input[text() or "x"].dim.extend([1, 3])
            acc = c.config.optimization.execution_accelerators.gpu_execution_accelerator.add()
            acc.name = "tensorrt"
            acc.parameters["precision_mode"] = rng.choice(["FP16", "FP32"])
            c.config.optimization.priority = rng.randrange(3)
        for _ in range(rng.randrange(2)):
            wu = c.config.model_warmup.add()
            wu.name, wu.batch_size = text(), rng.randrange(4)
            wi = wu.inputs[text() or "in"]
            wi.data_type = rng.randrange(15)
            wi.dims.extend([3])
            k = rng.randrange(3)
            if k == 0:
                wi.zero_data = True
            elif k == 1:
                wi.random_data = rng.random() < 0.5
            else:
                wi.input_data_file = text()
        if rng.random() < 0.3:
            c.config.response_cache.enable = True
        out.append(("ModelConfigResponse", c))
        md = service_pb2.ModelMetadataResponse(name=text(), platform=text())
        md.versions.extend(text(3) for _ in range(rng.randrange(3)))
        for _ in range(rng.randrange(3)):
            t = md.inputs.add()
            t.name, t.datatype = text(), "INT64"
            t.shape.extend([-1, 384])
        out.append(("ModelMetadataResponse", md))
        st = service_pb2.ModelStatisticsResponse()
        for _ in range(rng.randrange(3)):
            ms = st.model_stats.add()
            ms.name, ms.version, ms.inference_count = text(), "1", rng.randrange(10**12)
            ms.inference_stats.success.count = rng.randrange(10**6)
            ms.inference_stats.success.ns = rng.randrange(10**15)
            ms.inference_stats.queue.ns = rng.randrange(10**9)
            for _ in range(rng.randrange(3)):
                b = ms.batch_stats.add()
                b.batch_size = rng.randrange(64)
                b.compute_infer.count = rng.randrange(100)
            for _ in range(rng.randrange(2)):
                mu = ms.memory_usage.add()
                mu.type, mu.id, mu.byte_size = "GPU", rng.randrange(8), rng.randrange(2**40)
            for _ in range(rng.randrange(3)):
                ms.response_stats[str(rng.randrange(5))].success.count = rng.randrange(100)
        out.append(("ModelStatisticsResponse", st))
        cs = service_pb2.CudaSharedMemoryStatusResponse()
        for _ in range(rng.randrange(4)):
            name = text() or "r"
            cs.regions[name].name, cs.regions[name].device_id, cs.regions[name].byte_size = name, rng.randrange(8), rng.randrange(2**33)
        out.append(("CudaSharedMemoryStatusResponse", cs))
        ld = service_pb2.RepositoryModelLoadRequest(model_name=text())
        for _ in range(rng.randrange(3)):
            p = ld.parameters[text() or "file:x"]
            if rng.random() < 0.5:
                p.bytes_param = blob()
            else:
                p.string_param = text()
        out.append(("RepositoryModelLoadRequest", ld))
        lg = service_pb2.LogSettingsRequest()
        for _ in range(rng.randrange(3)):
            v = lg.settings[text() or "log_info"]
            k = rng.randrange(3)
            if k == 0:
                v.bool_param = rng.random() < 0.5
            elif k == 1:
                v.uint32_param = rng.choice([0, 2**32 - 1, 3])
            else:
                v.string_param = text()
        out.append(("LogSettingsRequest", lg))
    assert model_config_pb2.TYPE_FP32 == 11
    return out


def test_codec_round_trips_match_libprotobuf(binary, tmp_path):
    """Randomly filled messages serialised by libprotobuf (Python) are parsed and re-serialised by
    the C++ classes to the same bytes, and print the same text format."""
    from google.protobuf import text_format

    msgs = _random_messages(random.Random(20240921))
    path = tmp_path / "messages.txt"
    path.write_text("".join("%s %s\n" % (kind, m.SerializeToString(deterministic=True).hex() or "-") for kind, m in msgs))
    r = subprocess.run([binary, "--roundtrip", str(path)], capture_output=True, timeout=120)
    blocks = r.stdout.decode("utf-8", "surrogateescape").split("---\n")
    assert len(blocks) == len(msgs) + 1, r.stdout[-500:]
    text_mismatch = byte_mismatch = 0
    for (kind, m), block in zip(msgs, blocks):
        hex_line, _, text = block.partition("\n")
        want = m.SerializeToString(deterministic=True)
        if hex_line != want.hex():
            # only the order of map entries may differ (std::map orders keys bytewise, upb orders a
            # key after the keys it is a prefix of): same length, and the same message when parsed
            byte_mismatch += 1
            assert len(hex_line) == 2 * len(want) and type(m).FromString(bytes.fromhex(hex_line)) == m, (kind, m)
            continue
        expect = text_format.MessageToString(m, as_utf8=False)
        if text != expect:
            # the one tolerated difference: how the shortest float / double representation is printed
            has_float = any(tok in expect for tok in ("fp32_contents", "fp64_contents", "double_param"))
            assert has_float, (kind, text, expect)
            text_mismatch += 1
    assert text_mismatch < len(msgs) // 4 and byte_mismatch < len(msgs) // 10, (text_mismatch, byte_mismatch)


def test_loopback_against_grpcio_server(binary):
    proc, _, grpc_port = start_server()
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % grpc_port], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS (offline + loopback)" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
    finally:
        proc.terminate()
        proc.wait(10)


def test_client_timeout_is_deadline_exceeded(binary):
    proc, _, grpc_port = start_server(["--delay-us", "600000"])
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % grpc_port, "slow"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
    finally:
        proc.terminate()
        proc.wait(10)


def _has_gpu():
    import ctypes

    from client_b200 import _native

    n = ctypes.c_int(0)
    return _native.load().tb200_device_count(ctypes.byref(n)) == 0 and n.value > 0


def test_request_compression_needs_the_device(binary):
    if _has_gpu():
        pytest.skip("a GPU is present")
    proc, _, grpc_port = start_server()
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % grpc_port, "compress-nogpu"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
    finally:
        proc.terminate()
        proc.wait(10)


@pytest.mark.parametrize("algorithm", ["Gzip", "Deflate"])
def test_compressed_responses_are_inflated(binary, algorithm):
    """A grpcio server configured to compress its responses (message flag 1 + grpc-encoding): the
    client announces identity,deflate,gzip and inflates with zlib."""
    from concurrent import futures

    import grpc

    from client_b200.grpc import service_pb2, service_pb2_grpc

    seen = []

    class Identity(service_pb2_grpc.GRPCInferenceServiceServicer):
        def ModelInfer(self, request, context):
            seen.append(dict(context.invocation_metadata()).get("grpc-accept-encoding"))
            resp = service_pb2.ModelInferResponse(model_name=request.model_name, model_version="1")
            out = resp.outputs.add()
            out.name, out.datatype = "OUTPUT0", request.inputs[0].datatype
            out.shape.extend(request.inputs[0].shape)
            resp.raw_output_contents.append(request.raw_input_contents[0])
            return resp

    srv = grpc.server(futures.ThreadPoolExecutor(max_workers=2), compression=getattr(grpc.Compression, algorithm))
    service_pb2_grpc.add_GRPCInferenceServiceServicer_to_server(Identity(), srv)
    port = srv.add_insecure_port("127.0.0.1:0")
    srv.start()
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % port, "compressed-responses"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
        assert len(seen) == 3
    finally:
        srv.stop(0)


@pytest.mark.gpu
def test_request_compression_on_the_device(binary):
    """gzip / deflate message encoding produced by tb200_deflate_async, inflated by the grpcio server."""
    proc, _, grpc_port = start_server()
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % grpc_port, "compress-gpu"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
    finally:
        proc.terminate()
        proc.wait(10)


@pytest.mark.gpu
def test_reference_grpc_cudashm_example_if_prebuilt():
    """The reference's simple_grpc_cudashm_client.cc (cudaMalloc + cudaIpcGetMemHandle +
    RegisterCudaSharedMemory over gRPC), compiled unmodified in the build container
    (oracle/build_ref_examples.py -> oracle/_ref/), against the mock server's gRPC port."""
    exe = os.path.join(ROOT, "oracle", "_ref", "cc_examples", "simple_grpc_cudashm_client")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/cc_examples was not prebuilt")
    proc, _, grpc_port = start_server()
    try:
        r = subprocess.run([exe, "-u", "127.0.0.1:%d" % grpc_port], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS : Cuda Shared Memory" in r.stdout, r.stdout[-800:] + r.stderr[-400:]
    finally:
        proc.terminate()
        proc.wait(10)


def test_thread_sanitizer_clean(tmp_path):
    """The client sources under -fsanitize=thread: offline checks (incl. the reconnect and
    200-calls-in-flight cases against the in-process stub) and the loopback run report no race."""
    cpp = os.path.join(ROOT, "client_b200", "cpp")
    exe = str(tmp_path / "test_cc_grpc_tsan")
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-I" + os.path.join(cpp, "compat"), "-I" + cpp,
                            "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_cc_grpc_client.cc"),
                            os.path.join(cpp, "tb200_client.cc"), os.path.join(cpp, "tb200_grpc_client.cc"), "-o", exe,
                            "-L" + LIBDIR, "-ltb200", "-Wl,-rpath," + LIBDIR, "-lpthread", "-lz"], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime for g++ here: " + build.stderr[-200:])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0")
    first = ""
    for attempt in range(2):
        proc, _, grpc_port = start_server()
        try:
            r = subprocess.run([exe, "127.0.0.1:%d" % grpc_port], capture_output=True, text=True, timeout=300, env=env)
        finally:
            proc.terminate()
            proc.wait(10)
        out = r.stdout + r.stderr
        # a race report fails at once; a failed CHECK (the instrumented binary is ~10x slower, its
        # time limits are the loopback test's) gets one more run, and both outputs are shown
        assert "ThreadSanitizer" not in out, out[-3000:]
        if r.returncode == 0 and "PASS (offline + loopback)" in r.stdout:
            return
        first = first or out[-1500:]
    raise AssertionError("instrumented run failed twice:\n" + first + "\n---\n" + out[-1500:])
