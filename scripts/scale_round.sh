#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it (one rank per GPU; replicas, no data-path collective):
#   gpurun --gpus N -- bash scripts/scale_round.sh N [steps]
N=${1:-2}
STEPS=${2:-200}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 \
      bench.py --impl reference --gpus $N --steps $STEPS --warmup 5 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err
else
  python bench.py --impl reference --gpus $N --steps $STEPS --warmup 5 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err
fi
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
else
  python bench.py --gpus 1 --steps $STEPS --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
fi
tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_ref_n$N.json", "gpurun_out/r02_bench_n$N.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "fill_once", d.get("fill_once"), "roofline", d.get("roofline", {}).get("frac"), d.get("loopback_error"))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
