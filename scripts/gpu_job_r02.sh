cd "${GRAFT_REPO_ROOT:-.}"
python -m pytest tests/test_kernels_gpu.py -x -q -k "bytes or fill_homog or overlap" 2>&1 | tail -4
python -m pytest tests/test_reference_unit_tests.py tests/test_cudashm_gpu.py -x -q -m gpu 2>&1 | tail -4
python -m pytest tests/test_perf_gpu.py -x -q -k "pipelined" 2>&1 | tail -4
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps_pipe CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log
mkdir -p $CUDA_MPS_PIPE_DIRECTORY $CUDA_MPS_LOG_DIRECTORY
timeout 30 nvidia-cuda-mps-control -d; sleep 2
python scripts/h2d_compare.py 1 32 > gpurun_out/r02_h2d_compare.txt 2>&1
echo quit | timeout 30 nvidia-cuda-mps-control
unset CUDA_MPS_PIPE_DIRECTORY CUDA_MPS_LOG_DIRECTORY
cat gpurun_out/r02_h2d_compare.txt
bash scripts/ncu_round2.sh 2>&1 | tail -8
