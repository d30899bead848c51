set -u
mkdir -p gpurun_out
out=gpurun_out/r02d
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 12 "${out}_${name}.log" | grep -v Warning | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step gruseq_p3 200 env CTGCN_PAIR_VARIANT=3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "gru_seq_kernel and auto"
step ab       300 python profiles/gru_ab.py --impls unpaired,pair1,pair2,pair3
step parity   400 python -m pytest tests/test_parity_gpu.py tests/test_chunked_gpu.py -m gpu -x -q
step tl_p3  120 python profiles/gru_timeline.py --steps 6 --impl pair3
step tl_p3t 120 python profiles/gru_timeline.py --steps 6 --impl pair3 --mode 1
