set -u
mkdir -p gpurun_out
out=gpurun_out/r02u
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 6 "${out}_${name}.log" | grep -v Warning | cut -c1-1500 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step lin   200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "linear_kernel or mlp_golden or model_golden"
step cfg3  200 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline
