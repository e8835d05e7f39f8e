#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list and full captures.
# Everything is written under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "build rc=$?"
timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 100 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cat gpurun_out/bench_ref.json
if [ -z "$SKIP_NCU" ]; then
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-loopback > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 3 -c 2 -f -o gpurun_out/prof_fill \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-loopback > gpurun_out/ncu_fill.log 2>&1
echo "ncu fill rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:pack_image_chw -s 1 -c 2 -f -o gpurun_out/prof_pack \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-loopback > gpurun_out/ncu_pack.log 2>&1
echo "ncu pack rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none -k regex:check_kernel -s 1 -c 1 -f -o gpurun_out/prof_check \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-loopback > gpurun_out/ncu_check.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:resize_pack -s 2 -c 1 -f -o gpurun_out/prof_resize \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-loopback > gpurun_out/ncu_resize.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:fill_unaligned -s 3 -c 1 -f -o gpurun_out/prof_fill_unaligned \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-loopback > gpurun_out/ncu_fill_unaligned.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:deflate_chunk -s 2 -c 1 -f -o gpurun_out/prof_deflate \
    python scripts/deflate_probe.py > gpurun_out/ncu_deflate.log 2>&1
fi
ls -la gpurun_out
