# round-2 validation on one B200: the whole -m gpu suite, smoke(), both bench arms
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_gpu_tests.log 2>&1
tail -5 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -c 600 gpurun_out/r02_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_b200.json 2> gpurun_out/r02_bench_b200.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench_b200.json").read().strip().splitlines()[-1])
for k in ("value", "e2e", "fill_once", "roofline", "deflate", "resize_pack", "clocks", "loopback_error"):
    print(k, json.dumps(d.get(k))[:700])
P
tail -5 gpurun_out/r02_bench_b200.err
