set -u
mkdir -p gpurun_out
out=gpurun_out/r02t
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 2 "${out}_${name}.log" | grep -v Warning | cut -c1-1500 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step cfg3  200 python bench.py --config cfg3 --steps 10 --warmup 3
step cfg5s 300 python bench.py --config cfg5s --steps 3 --warmup 3 --no-cpu-baseline
step cfg2  200 python bench.py --config cfg2 --steps 20 --warmup 3
