cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_cudashm_gpu.py -x -q -k "resize or image" 2>&1 | tail -4
python scripts/resize_probe.py 50 2>&1 | tail -4
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:resize_pack -s 6 -c 1 -f -o gpurun_out/r02_prof_resize python scripts/resize_probe.py 2 > gpurun_out/ncu_resize.log 2>&1
echo "ncu rc=$?"
