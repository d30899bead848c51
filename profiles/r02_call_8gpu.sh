set -u
mkdir -p gpurun_out
out=gpurun_out/r02v
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; grep '^{' "${out}_${name}.log" | tail -n 1 | cut -c1-4000 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
step copy      120 $TR --master-port 29521 profiles/host_copy_ceiling.py --mb 512
step copy_nb   120 $TR --master-port 29522 profiles/host_copy_ceiling.py --mb 512 --no-numa-bind
step bench8    240 $TR --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3
nvidia-smi topo -m > "${out}_topo.log" 2>&1; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> "${out}_topo.log" 2>&1
