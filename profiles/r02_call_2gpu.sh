set -u
mkdir -p gpurun_out
out=gpurun_out/r02s
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 3 "${out}_${name}.log" | grep -v Warning | cut -c1-3000 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step dist_tests 300 python -m pytest tests/test_dist_gpu.py -m gpu -x -q
step bench2 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
