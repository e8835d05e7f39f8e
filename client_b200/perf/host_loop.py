"""The C2 request loop a tritonclient user writes, on the drop-in modules: host numpy tensor ->
``cuda_shared_memory.set_shared_memory_region`` (host->device) -> ``InferenceServerClient.infer``
naming the regions -> ``get_contents_as_numpy`` (device->host).  This is the flow of the
reference's src/python/examples/simple_http_cudashm_client.py:82-139 with ``client_b200`` in the
place of ``tritonclient``; bench.py times it (``e2e``) beside the same loop on the restated
reference code (oracle/ref_client.py, ``--impl reference``).  The device-side generator
(``client_b200.perf`` native engine) is the path with no host tensors at all."""

import time

import numpy as np

IN_SHAPE = (3, 224, 224)
IN_BYTES = 3 * 224 * 224 * 4
OUT_ELEMS = 1000
OUT_BYTES = OUT_ELEMS * 4


def run_loop(url, device_id, tag, seconds, data_mode, ready=None, go=None):
    """Free-running closed loop of ONE client for ``seconds``; returns (completed, latencies_ns).
    ``data_mode``: "per-request" (fresh host tensor, H2D, infer, D2H every request), "device"
    (every request's tensor generated inside the region by ``fill_shared_memory_region``, infer, the
    response validated by ``check_shared_memory_region`` -- only the job descriptors and the verdict
    cross PCIe; the check of one response and the fill for the next request share one wait) or "once" (regions filled once, every request only names them)."""
    from .. import http as httpclient
    from ..utils import cuda_shared_memory as cudashm

    client = httpclient.InferenceServerClient(url)
    in_name, out_name = "host_in_%s" % tag, "host_out_%s" % tag
    in_h = cudashm.create_shared_memory_region(in_name, IN_BYTES, device_id)
    out_h = cudashm.create_shared_memory_region(out_name, OUT_BYTES, device_id)
    client.register_cuda_shared_memory(in_name, cudashm.get_raw_handle(in_h), device_id, IN_BYTES)
    client.register_cuda_shared_memory(out_name, cudashm.get_raw_handle(out_h), device_id, OUT_BYTES)
    inp = httpclient.InferInput("data_0", list(IN_SHAPE), "FP32").set_shared_memory(in_name, IN_BYTES)
    out = httpclient.InferRequestedOutput("fc6_1")
    out.set_shared_memory(out_name, OUT_BYTES)
    rng = np.random.default_rng(abs(hash(tag)) % (1 << 32))
    cudashm.set_shared_memory_region(in_h, [rng.random(IN_SHAPE, dtype=np.float32)])
    client.infer("densenet_onnx", [inp], outputs=[out])
    if data_mode == "device":
        cudashm.fill_shared_memory_region(in_h, "FP32", IN_SHAPE, seed=1)  # the first request's tensor
    if ready is not None:
        ready.wait()
    if go is not None:
        go.wait()
    lat = []
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        t0 = time.perf_counter_ns()
        if data_mode == "per-request":
            cudashm.set_shared_memory_region(in_h, [rng.random(IN_SHAPE, dtype=np.float32)])
            client.infer("densenet_onnx", [inp], outputs=[out])
            y = cudashm.get_contents_as_numpy(out_h, np.float32, [OUT_ELEMS])
            if not np.isfinite(y).all():
                raise RuntimeError("non-finite logits")
        elif data_mode == "device":
            # the region already holds this request's tensor (generated while the previous response was checked)
            client.infer("densenet_onnx", [inp], outputs=[out])
            verdict = cudashm.check_shared_memory_region(out_h, "top1", byte_size=OUT_BYTES, defer=True)
            cudashm.fill_shared_memory_region(in_h, "FP32", IN_SHAPE, seed=len(lat) + 2, sync=False)  # the next request's tensor
            if verdict()["mismatches"]:  # one wait for both kernels; mismatches = non-finite values
                raise RuntimeError("non-finite logits")
        else:
            client.infer("densenet_onnx", [inp], outputs=[out])
        lat.append(time.perf_counter_ns() - t0)
    client.unregister_cuda_shared_memory(in_name)
    client.unregister_cuda_shared_memory(out_name)
    cudashm.destroy_shared_memory_region(in_h)
    cudashm.destroy_shared_memory_region(out_h)
    client.close()
    return len(lat), lat
