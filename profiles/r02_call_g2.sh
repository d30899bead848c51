set -u
mkdir -p gpurun_out
out=gpurun_out/r02S
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 8 "${out}_${name}.log" | grep -v Warning | cut -c1-400 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step parity 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "gru_seq or core_diffusion or cdn or model_golden"
step ab128 120 python profiles/gru_ab.py --n 1000000 --impls unpaired,auto --iters 10
step ab96 120 python profiles/gru_ab.py --n 1000000 --d-in 96 --impls auto --iters 5
