cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_cudashm_gpu.py tests/test_native_server_gpu.py -x -q -k "resize or image or host_loop" 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_b200.json 2> gpurun_out/r02_bench_b200.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench_b200.json").read().strip().splitlines()[-1])
for k in ("value", "e2e", "fill_once", "resize_pack", "clocks", "loopback_error"):
    print(k, json.dumps(d.get(k))[:900])
P
tail -5 gpurun_out/r02_bench_b200.err
