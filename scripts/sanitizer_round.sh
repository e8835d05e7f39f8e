#!/bin/bash
# compute-sanitizer passes over the GPU kernel tests (memcheck on everything small enough,
# racecheck / synccheck on the shared-memory kernels).  Summaries -> profiles/rNN_sanitizer.txt
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_kernels_gpu.py tests/test_cudashm_gpu.py -q -p no:cacheprovider -k "not 4000 and not 1080" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"
for tool in racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_kernels_gpu.py tests/test_cudashm_gpu.py -q -p no:cacheprovider -k "(pack or resize or topk or check or classify or deflate_round) and not 4000 and not 1080 and not 600" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
done
