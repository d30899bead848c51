set -u
mkdir -p gpurun_out
out=gpurun_out/r02U
free -g > "${out}_mem.log"; nvidia-smi --query-gpu=memory.total --format=csv >> "${out}_mem.log"
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
echo "MemAvailable ${avail} GB" | tee -a "${out}_summary.log"
if [ "$avail" -lt 350 ]; then echo "not enough host memory for the pinned e2e buffers of cfg5 (8 x 30 GB): skipped" | tee -a "${out}_summary.log"; exit 0; fi
t0=$SECONDS
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > "${out}_cfg5.log" 2>&1
echo "[cfg5 x8] rc=$? $((SECONDS - t0))s" | tee -a "${out}_summary.log"
grep '^{' "${out}_cfg5.log" | tail -n 1 | cut -c1-6000 | tee -a "${out}_summary.log"
tail -n 5 "${out}_cfg5.log" | cut -c1-400 >> "${out}_summary.log"
