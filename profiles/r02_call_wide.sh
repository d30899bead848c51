set -u
mkdir -p gpurun_out
out=gpurun_out/r02w
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 12 "${out}_${name}.log" | cut -c1-400 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step gruseq 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "gru_seq and (wide or auto)"
step parity_wide 400 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wide or (auto and 256)"
step ab256 200 python profiles/gru_ab.py --n 400000 --h 256 --d-in 256 --impls wide,simt --iters 3
for u in 2 3 6 8; do CTGCN_WIDE_UNITS=$u step ab256_u$u 120 python profiles/gru_ab.py --n 400000 --h 256 --d-in 256 --impls wide --iters 5; done
step ab128 120 python profiles/gru_ab.py --n 1000000 --impls auto,wide --iters 5
step ab512 120 python profiles/gru_ab.py --n 200000 --h 512 --d-in 512 --impls wide --iters 3
