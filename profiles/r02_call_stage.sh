set -u
mkdir -p gpurun_out
out=gpurun_out/r02K
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 9 "${out}_${name}.log" | grep -v Warning | cut -c1-400 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step parity 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu
step ab128 120 python profiles/gru_ab.py --n 1000000 --impls one_cta_r1,unpaired,auto --iters 10
step ab500 120 python profiles/gru_ab.py --n 300000 --d-in 500 --impls auto --iters 5
step ab256 120 python profiles/gru_ab.py --n 400000 --h 256 --d-in 256 --impls wide --iters 5
step linab 120 python profiles/linear_ab.py
step cfg4 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
