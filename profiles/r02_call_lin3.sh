set -u
mkdir -p gpurun_out
out=gpurun_out/r02D
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 10 "${out}_${name}.log" | cut -c1-300 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step tests 400 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "linear or mlp or gru_seq or (model_golden and auto) or 256"
step linab 200 python profiles/linear_ab.py
step ab256 200 python profiles/gru_ab.py --n 400000 --h 256 --d-in 256 --impls wide --iters 5
step tl256 120 python profiles/gru_wide_timeline.py --h 256 --d-in 256 --steps 4
step tl256_m1 120 python profiles/gru_wide_timeline.py --h 256 --d-in 256 --steps 4 --mode 1
