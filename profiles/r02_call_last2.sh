set -u
mkdir -p gpurun_out
out=gpurun_out/r02Z
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; }
step cfg3 100 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline
step cfg2 100 python bench.py --config cfg2 --steps 20 --warmup 3 --no-cpu-baseline
