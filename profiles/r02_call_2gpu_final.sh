set -u
mkdir -p gpurun_out
out=gpurun_out/r02W
t0=$SECONDS
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > "${out}_bench2.log" 2>&1
echo "[bench2] rc=$? $((SECONDS - t0))s" | tee -a "${out}_summary.log"
grep '^{' "${out}_bench2.log" | tail -n 1 | cut -c1-3000 | tee -a "${out}_summary.log"
