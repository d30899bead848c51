set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_train_gpu.py -x -q -m gpu > gpurun_out/r02AA_train.log 2>&1; echo "[train] rc=$?" | tee gpurun_out/r02AA_summary.log; tail -n 3 gpurun_out/r02AA_train.log | tee -a gpurun_out/r02AA_summary.log
