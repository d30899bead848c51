set -u
mkdir -p gpurun_out
out=gpurun_out/r02P
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 3 "${out}_${name}.log" | grep -v Warning | cut -c1-3000 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step bench_default 600 python bench.py
step bench_reference 900 python bench.py --impl reference
