# resize_pack_kernel: CUDA-event times for three destination types, then one ncu --set full capture
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python scripts/resize_probe.py 50 2>&1 | tail -4 | tee gpurun_out/r02_resize_probe.txt
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:resize_pack -s 6 -c 1 -f -o gpurun_out/r02_prof_resize python scripts/resize_probe.py 2 > gpurun_out/ncu_resize.log 2>&1
echo "ncu rc=$?"
