set -u
mkdir -p gpurun_out
out=gpurun_out/r02G
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 6 "${out}_${name}.log" | cut -c1-600 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step hubtests 300 python -m pytest tests/test_parity_gpu.py tests/test_chunked_gpu.py tests/test_train_gpu.py -x -q -m gpu -k "hub or cumspmm or chunked or plan or powerlaw"
step cfg5s 400 python bench.py --config cfg5s --steps 3 --warmup 3
free -g | tee -a "${out}_summary.log"
