set -u
mkdir -p gpurun_out
out=gpurun_out/r02X
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 9 "${out}_${name}.log" | grep -v Warning | cut -c1-300 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step lintests 200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "linear or mlp or (model_golden and auto)"
step linab 100 python profiles/linear_ab.py
