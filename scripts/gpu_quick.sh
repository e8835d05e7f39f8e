#!/bin/bash
# Short GPU call: selected tests + sweeps (no ncu).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python scripts/fill_sweep.py > gpurun_out/fill_sweep.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/fill_sweep.txt
