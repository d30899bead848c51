set -u
mkdir -p gpurun_out
out=gpurun_out/r02T
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 4 "${out}_${name}.log" | grep -v Warning | cut -c1-900 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step alltests 900 python -m pytest tests -x -q -m gpu
step cfg4 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
step cfg5s 300 python bench.py --config cfg5s --steps 3 --warmup 3 --no-cpu-baseline
