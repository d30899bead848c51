set -u
mkdir -p gpurun_out
out=gpurun_out/r02O
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 7 "${out}_${name}.log" | grep -v Warning | cut -c1-700 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step alltests 900 python -m pytest tests -x -q -m gpu
step ab256 120 python profiles/gru_ab.py --n 400000 --h 256 --d-in 256 --impls unpaired,wide --iters 5
step cfg5s 300 python bench.py --config cfg5s --steps 3 --warmup 3 --no-cpu-baseline
step cfg3 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline
step cfg2 300 python bench.py --config cfg2 --steps 20 --warmup 3
