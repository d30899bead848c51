set -u
mkdir -p gpurun_out
out=gpurun_out/r02Y
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 12 "${out}_${name}.log" | grep -v Warning | cut -c1-400 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step chunks 120 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "row_chunks"
step alltests 400 python -m pytest tests -x -q -m gpu
