"""The native load generator's transport on the CPU box: closed loop against the
canned-response server and against the Python mock server (system shm, no device work)."""

import ctypes

import numpy as np

from client_b200 import _native
from client_b200._native_loadgen import LoadgenConfig, LoadgenStats
from client_b200.perf.native import StubServer, frame_http_request
from test_loopback import start_server


def _run(host, port, reqs, seconds=0.5):
    lib = _native.load()
    n = len(reqs)
    bufs = [ctypes.create_string_buffer(r, len(r)) for r in reqs]
    cfg = LoadgenConfig()
    cfg.host, cfg.port, cfg.concurrency = host.encode(), port, n
    cfg.requests = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
    cfg.request_sizes = (ctypes.c_uint64 * n)(*[len(r) for r in reqs])
    h = ctypes.c_void_p()
    _native.check(lib.tb200_loadgen_create(ctypes.byref(cfg), ctypes.byref(h)))
    _native.check(lib.tb200_loadgen_start(h))
    st = LoadgenStats()
    _native.check(lib.tb200_loadgen_window(h, 0.2, ctypes.byref(st)))  # warm-up window
    _native.check(lib.tb200_loadgen_window(h, seconds, ctypes.byref(st)))
    lib.tb200_loadgen_stop(h)
    lib.tb200_loadgen_destroy(h)
    return st


def test_closed_loop_against_stub_server():
    srv = StubServer()
    try:
        body = b'{"inputs":[{"name":"INPUT0","shape":[1,16],"datatype":"INT32","parameters":{"shared_memory_region":"r","shared_memory_byte_size":64}}]}'
        req = frame_http_request(srv.host, srv.port, "v2/models/simple/infer", body, None)
        st = _run(srv.host, srv.port, [req] * 4)
        assert st.failed_request_count == 0 and st.completed_request_count > 200
        assert 0 < st.min_ns <= st.p50_ns <= st.p99_ns <= st.max_ns
        avg = st.cumulative_total_request_time_ns / st.completed_request_count
        assert st.cumulative_send_time_ns + st.cumulative_receive_time_ns <= st.cumulative_total_request_time_ns
        assert abs(st.completed_request_count / st.window_seconds - 4 / (avg * 1e-9)) / (4 / (avg * 1e-9)) < 0.5  # Little's law, roughly
    finally:
        srv.stop()


def test_closed_loop_against_mock_server_system_shm():
    """Requests that only name system-shm regions; the mock server computes add/sub."""
    import client_b200.http as httpclient
    import client_b200.utils.shared_memory as shm

    proc, http_port, _ = start_server()
    try:
        a = np.arange(16, dtype=np.int32)
        b = np.ones(16, dtype=np.int32)
        ip = shm.create_shared_memory_region("lg_in", "/tb200_lg_in", 128)
        op = shm.create_shared_memory_region("lg_out", "/tb200_lg_out", 128)
        shm.set_shared_memory_region(ip, [a, b])
        with httpclient.InferenceServerClient("127.0.0.1:%d" % http_port) as client:
            client.register_system_shared_memory("lg_in", "/tb200_lg_in", 128)
            client.register_system_shared_memory("lg_out", "/tb200_lg_out", 128)
            inputs = [httpclient.InferInput("INPUT0", [1, 16], "INT32").set_shared_memory("lg_in", 64),
                      httpclient.InferInput("INPUT1", [1, 16], "INT32").set_shared_memory("lg_in", 64, offset=64)]
            outputs = [httpclient.InferRequestedOutput("OUTPUT0"), httpclient.InferRequestedOutput("OUTPUT1")]
            outputs[0].set_shared_memory("lg_out", 64)
            outputs[1].set_shared_memory("lg_out", 64, offset=64)
            body, js = client.generate_request_body(inputs, outputs=outputs)
            req = frame_http_request("127.0.0.1", http_port, "v2/models/simple/infer", body, js)
            st = _run("127.0.0.1", http_port, [req, req], seconds=0.5)
            assert st.failed_request_count == 0 and st.completed_request_count > 10
            got = shm.get_contents_as_numpy(op, np.int32, [2, 16])
            assert np.array_equal(got[0], a + b) and np.array_equal(got[1], a - b)
            del got
            bad = frame_http_request("127.0.0.1", http_port, "v2/models/nope/infer", body, js)
            st = _run("127.0.0.1", http_port, [bad], seconds=0.2)
            assert st.completed_request_count == 0 and st.failed_request_count > 0
            client.unregister_system_shared_memory()
        shm.destroy_shared_memory_region(ip)
        shm.destroy_shared_memory_region(op)
    finally:
        proc.terminate()
        proc.wait(10)


def test_many_connections_per_thread_and_large_borrowed_tails():
    """40 slots over a handful of event-loop threads, every request followed by a 1 MiB
    borrowed tail (HTTP binary-tensor body in --shared-memory none mode): the non-blocking
    send has to continue across EPOLLOUT rounds, and the server must see whole bodies."""
    srv = StubServer()
    lib = _native.load()
    try:
        n, tail_bytes = 40, 1 << 20
        head = b'{"inputs":[{"name":"X","shape":[%d],"datatype":"UINT8","parameters":{"binary_data_size":%d}}]}' % (tail_bytes, tail_bytes)
        req = frame_http_request(srv.host, srv.port, "v2/models/m/infer", head, len(head), head_only_bytes=tail_bytes)
        tail = ctypes.create_string_buffer(bytes(range(256)) * (tail_bytes // 256), tail_bytes)
        bufs = [ctypes.create_string_buffer(req, len(req)) for _ in range(n)]
        cfg = LoadgenConfig()
        cfg.host, cfg.port, cfg.concurrency = srv.host.encode(), srv.port, n
        cfg.requests = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
        cfg.request_sizes = (ctypes.c_uint64 * n)(*[len(req)] * n)
        cfg.tails = (ctypes.c_void_p * n)(*[ctypes.addressof(tail)] * n)
        cfg.tail_sizes = (ctypes.c_uint64 * n)(*[tail_bytes] * n)
        h = ctypes.c_void_p()
        _native.check(lib.tb200_loadgen_create(ctypes.byref(cfg), ctypes.byref(h)))
        _native.check(lib.tb200_loadgen_start(h))
        st = LoadgenStats()
        _native.check(lib.tb200_loadgen_window(h, 1.0, ctypes.byref(st)))
        lib.tb200_loadgen_stop(h)
        lib.tb200_loadgen_destroy(h)
        assert st.failed_request_count == 0 and st.completed_request_count >= n, (st.completed_request_count, st.failed_request_count)
        assert st.cumulative_send_time_ns > 0
    finally:
        srv.stop()


# ---- gRPC over cleartext HTTP/2 (csrc/h2.h) -----------------------------------------------
def _grpc_request(model, arrays, with_data=True):
    import client_b200.grpc as grpcclient
    from client_b200.grpc._utils import _get_inference_request

    ins = []
    for name, arr in arrays:
        inp = grpcclient.InferInput(name, list(arr.shape), "INT32")
        if with_data:
            inp.set_data_from_numpy(arr)
        ins.append(inp)
    return _get_inference_request(model_name=model, inputs=ins, model_version="", request_id="", outputs=None, sequence_id=0,
                                  sequence_start=False, sequence_end=False, priority=0, timeout=None, parameters=None).SerializeToString()


def _run_grpc(host, port, reqs, tails=None, seconds=0.4):
    lib = _native.load()
    n = len(reqs)
    bufs = [ctypes.create_string_buffer(r, len(r)) for r in reqs]
    cfg = LoadgenConfig()
    cfg.host, cfg.port, cfg.concurrency, cfg.protocol = host.encode(), port, n, 1
    cfg.requests = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
    cfg.request_sizes = (ctypes.c_uint64 * n)(*[len(r) for r in reqs])
    keep = []
    if tails:
        keep = [ctypes.create_string_buffer(t, len(t)) for t in tails]
        cfg.tails = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in keep])
        cfg.tail_sizes = (ctypes.c_uint64 * n)(*[len(t) for t in tails])
    h = ctypes.c_void_p()
    _native.check(lib.tb200_loadgen_create(ctypes.byref(cfg), ctypes.byref(h)))
    _native.check(lib.tb200_loadgen_start(h))
    st = LoadgenStats()
    _native.check(lib.tb200_loadgen_window(h, seconds, ctypes.byref(st)))
    lib.tb200_loadgen_stop(h)
    lib.tb200_loadgen_destroy(h)
    return st


def test_grpc_transport_against_grpcio_server():
    """The native HTTP/2 client speaks to a real grpcio server (the mock server's gRPC port):
    message = protobuf head + borrowed tail (raw_input_contents tag/length/tensor), errors
    arrive as trailers-only responses, a 48 KB message spans several DATA frames."""
    a = np.arange(16, dtype=np.int32)[None, :]
    b = np.ones((1, 16), dtype=np.int32)
    full = _grpc_request("simple", [("INPUT0", a), ("INPUT1", b)])
    head = _grpc_request("simple", [("INPUT0", a), ("INPUT1", b)], with_data=False)
    tail = b"\x3a\x40" + a.tobytes() + b"\x3a\x40" + b.tobytes()
    assert head + tail == full  # raw_input_contents is the last field of ModelInferRequest
    proc, _, grpc_port = start_server()
    try:
        st = _run_grpc("127.0.0.1", grpc_port, [head] * 3, [tail] * 3)
        assert st.failed_request_count == 0 and st.completed_request_count > 20
        assert 0 < st.min_ns <= st.p50_ns <= st.max_ns
        st = _run_grpc("127.0.0.1", grpc_port, [_grpc_request("nope", [("INPUT0", a), ("INPUT1", b)])], seconds=0.2)
        assert st.completed_request_count == 0 and st.failed_request_count > 0
        big = np.arange(12000, dtype=np.int32)
        st = _run_grpc("127.0.0.1", grpc_port, [_grpc_request("custom_identity_int32", [("INPUT0", big)])] * 2, seconds=0.3)
        assert st.failed_request_count == 0 and st.completed_request_count > 5
    finally:
        proc.terminate()
        proc.wait(10)


def test_grpc_stub_server_with_grpcio_client_and_native_client():
    """The canned-response gRPC stub (HTTP/2 server side) is accepted by a real grpcio client,
    and the native client sustains a closed loop against it."""
    import client_b200.grpc as grpcclient
    from client_b200.grpc import service_pb2
    from client_b200.perf.native import GrpcStubServer

    resp = service_pb2.ModelInferResponse(model_name="stub", model_version="1")
    out = resp.outputs.add()
    out.name, out.datatype = "OUTPUT0", "INT32"
    out.shape.extend([1, 4])
    resp.raw_output_contents.append(np.arange(4, dtype=np.int32).tobytes())
    stub = GrpcStubServer(resp.SerializeToString())
    try:
        with grpcclient.InferenceServerClient(stub.url) as client:
            inp = grpcclient.InferInput("INPUT0", [1, 4], "INT32").set_data_from_numpy(np.ones((1, 4), np.int32))
            for _ in range(20):
                assert np.array_equal(client.infer("anything", [inp]).as_numpy("OUTPUT0"), np.arange(4, dtype=np.int32)[None, :])
        a = np.arange(16, dtype=np.int32)[None, :]
        st = _run_grpc(stub.host, stub.port, [_grpc_request("simple", [("INPUT0", a), ("INPUT1", a)])] * 8)
        assert st.failed_request_count == 0 and st.completed_request_count > 1000
    finally:
        stub.stop()


def test_unary_grpc_response_larger_than_the_stream_window():
    """a unary ModelInfer response of 3 MB against the 1 MiB stream window the client
    advertises -- the client has to hand stream-level credit back (WINDOW_UPDATE on the call's
    stream), or the server stalls on the second MiB and the slot never completes."""
    from client_b200.grpc import service_pb2
    from client_b200.perf.native import GrpcStubServer

    resp = service_pb2.ModelInferResponse(model_name="stub", model_version="1")
    out = resp.outputs.add()
    out.name, out.datatype = "OUTPUT0", "UINT8"
    out.shape.extend([3 << 20])
    resp.raw_output_contents.append(bytes(3 << 20))
    stub = GrpcStubServer(resp.SerializeToString())
    try:
        a = np.arange(16, dtype=np.int32)[None, :]
        st = _run_grpc(stub.host, stub.port, [_grpc_request("simple", [("INPUT0", a), ("INPUT1", a)])] * 2, seconds=0.6)
        assert st.failed_request_count == 0 and st.completed_request_count >= 4, (st.completed_request_count, st.failed_request_count)
    finally:
        stub.stop()


def _run_stream(host, port, reqs, seconds=0.4):
    lib = _native.load()
    n = len(reqs)
    bufs = [ctypes.create_string_buffer(r, len(r)) for r in reqs]
    cfg = LoadgenConfig()
    cfg.host, cfg.port, cfg.concurrency, cfg.protocol = host.encode(), port, n, 2
    cfg.requests = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
    cfg.request_sizes = (ctypes.c_uint64 * n)(*[len(r) for r in reqs])
    h = ctypes.c_void_p()
    _native.check(lib.tb200_loadgen_create(ctypes.byref(cfg), ctypes.byref(h)))
    _native.check(lib.tb200_loadgen_start(h))
    st = LoadgenStats()
    _native.check(lib.tb200_loadgen_window(h, seconds, ctypes.byref(st)))
    lib.tb200_loadgen_stop(h)
    lib.tb200_loadgen_destroy(h)
    return st


def test_grpc_stream_mode_against_grpcio_server():
    """protocol 2: one ModelStreamInfer stream per connection (reference start_stream /
    async_stream_infer).  Against the grpcio mock server: a decoupled model answers
    max_tokens times per request and the request completes at the response flagged
    triton_final_response; a plain model answers once; an unknown model fails the request
    through error_message and the stream stays usable."""
    import client_b200.grpc as grpcclient
    from client_b200.grpc._utils import _get_inference_request

    ins = [grpcclient.InferInput("input_ids", [1, 64], "INT32").set_data_from_numpy(np.arange(64, dtype=np.int32)[None, :])]
    llama = _get_inference_request(model_name="llama3_8b", inputs=ins, model_version="", request_id="", outputs=None, sequence_id=0,
                                   sequence_start=False, sequence_end=False, priority=0, timeout=None,
                                   parameters={"max_tokens": 5}).SerializeToString()
    z = np.zeros((1, 16), np.int32)
    proc, _, grpc_port = start_server()
    try:
        st = _run_stream("127.0.0.1", grpc_port, [llama] * 3)
        assert st.failed_request_count == 0 and st.completed_request_count > 10
        assert st.response_count == 5 * st.completed_request_count
        assert 0 < st.first_response_p50_ns <= st.p50_ns and st.first_response_p50_ns <= st.first_response_p99_ns
        st = _run_stream("127.0.0.1", grpc_port, [_grpc_request("simple", [("INPUT0", z), ("INPUT1", z)])] * 2, 0.3)
        assert st.failed_request_count == 0 and st.completed_request_count > 10 and st.response_count == st.completed_request_count
        st = _run_stream("127.0.0.1", grpc_port, [_grpc_request("nope", [("INPUT0", z)])], 0.2)
        assert st.completed_request_count == 0 and st.failed_request_count > 3
    finally:
        proc.terminate()
        proc.wait(10)


def test_grpc_stream_stub_with_grpcio_and_native_clients():
    """The streaming mode of the canned-response gRPC stub: N responses per request message,
    accepted by a real grpcio bidirectional stream; the native generator sustains the loop
    (requests larger than one HTTP/2 frame, many more bytes than one stream window)."""
    import threading

    import client_b200.grpc as grpcclient
    from client_b200.perf.native import GrpcStubServer, stream_token_responses

    resp, fin = stream_token_responses()
    stub = GrpcStubServer(resp, final_response=fin, responses_per_request=16)
    a = np.arange(4096, dtype=np.int32)[None, :]
    try:
        got, done = [], threading.Event()

        def on_response(result, error):
            got.append((result, error))
            if len(got) == 48:
                done.set()

        with grpcclient.InferenceServerClient(stub.url) as client:
            client.start_stream(callback=on_response)
            inp = grpcclient.InferInput("input_ids", [1, 4096], "INT32").set_data_from_numpy(a)
            for _ in range(3):
                client.async_stream_infer("anything", [inp])
            assert done.wait(10)
            client.stop_stream()
        assert all(e is None for _, e in got)
        finals = [r.get_response().parameters["triton_final_response"].bool_param for r, _ in got]
        assert finals == ([False] * 15 + [True]) * 3
        st = _run_stream(stub.host, stub.port, [_grpc_request("llama3_8b", [("input_ids", a)])] * 8, 0.5)
        assert st.failed_request_count == 0 and st.completed_request_count > 2000  # > 32 MB on each stream
        assert st.response_count == 16 * st.completed_request_count
    finally:
        stub.stop()


def test_grpc_server_core_with_real_clients():
    """csrc/grpc_server.h (the event-loop core of the native model server) behind the echo handler:
    a grpcio client's header blocks (incremental indexing, dynamic table) decode to the right :path,
    2 MB messages pass both flow-control directions, 20 messages share one bidirectional stream,
    and the native generator drives it in unary and stream mode."""
    import threading

    import client_b200.grpc as grpcclient
    from client_b200.perf.native import GrpcEchoServer

    srv = GrpcEchoServer()
    try:
        with grpcclient.InferenceServerClient(srv.url) as client:
            for n in (16, 20000, 500000):
                a = np.arange(n, dtype=np.int32)
                inp = grpcclient.InferInput("INPUT0", [n], "INT32").set_data_from_numpy(a)
                resp = client.infer("echo_model", [inp], outputs=[grpcclient.InferRequestedOutput("whatever")], request_id="id-%d" % n).get_response()
                # the request read back as a response: inputs -> outputs, requested outputs -> raw_output_contents
                assert (resp.model_name, resp.id, [o.name for o in resp.outputs]) == ("echo_model", "id-%d" % n, ["INPUT0"])
                assert list(resp.outputs[0].shape) == [n] and len(resp.raw_output_contents) == 1
            got, done = [], threading.Event()

            def on_response(result, error):
                got.append((result, error))
                if len(got) == 20:
                    done.set()

            client.start_stream(callback=on_response)
            big = grpcclient.InferInput("INPUT0", [100000], "INT32").set_data_from_numpy(np.arange(100000, dtype=np.int32))
            for i in range(20):
                client.async_stream_infer("m%d" % i, [big])
            assert done.wait(30)
            client.stop_stream()
            assert all(e is None for _, e in got)
            assert [r.get_response().model_name for r, _ in got] == ["m%d" % i for i in range(20)]
        z = np.zeros((1, 16), np.int32)
        small = _grpc_request("simple", [("INPUT0", z), ("INPUT1", z)])
        st = _run_grpc(srv.host, srv.port, [small] * 8)
        assert st.failed_request_count == 0 and st.completed_request_count > 1000
        st = _run_grpc(srv.host, srv.port, [_grpc_request("x", [("INPUT0", np.arange(12000, dtype=np.int32))])] * 4, seconds=0.3)
        assert st.failed_request_count == 0 and st.completed_request_count > 100
        st = _run_stream(srv.host, srv.port, [small] * 8)
        assert st.failed_request_count == 0 and st.completed_request_count > 1000 and st.response_count == st.completed_request_count
    finally:
        srv.stop()
