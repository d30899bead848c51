set -u
mkdir -p gpurun_out
out=gpurun_out/r02J
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 10 "${out}_${name}.log" | grep -v Warning | cut -c1-300 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step lintests 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "linear or mlp"
step linab 200 python profiles/linear_ab.py
step ncu_lin 200 ncu --set full --clock-control none -k regex:"linear_tc_kernel" -s 4 -c 1 -f -o gpurun_out/r02_prof_linear2 python profiles/linear_ab.py
