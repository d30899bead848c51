set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_loss_gpu.py tests/test_train_gpu.py -x -q -m gpu > gpurun_out/r02AB_loss.log 2>&1; echo "[loss+train] rc=$?" | tee gpurun_out/r02AB_summary.log; tail -n 2 gpurun_out/r02AB_loss.log | tee -a gpurun_out/r02AB_summary.log
