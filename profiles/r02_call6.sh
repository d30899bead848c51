set -u
mkdir -p gpurun_out
out=gpurun_out/r02f
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 6 "${out}_${name}.log" | grep -v Warning | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step tests    600 python -m pytest tests -m gpu -x -q
step bench    300 python bench.py --steps 10 --warmup 3
step cfg2     200 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline
step hubsplit 150 python profiles/try_hubsplit.py
