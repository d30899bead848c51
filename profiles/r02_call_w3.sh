set -u
mkdir -p gpurun_out
out=gpurun_out/r02R
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 14 "${out}_${name}.log" | grep -v Warning | cut -c1-400 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step gruseq 200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "gru_seq and (wide or auto or unpaired)"
step ab256 120 python profiles/gru_ab.py --n 400000 --h 256 --d-in 256 --impls unpaired,wide --iters 5
step tl256 100 python profiles/gru_wide_timeline.py --h 256 --d-in 256 --steps 4
