set -u
mkdir -p gpurun_out
out=gpurun_out/r02b
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 8 "${out}_${name}.log" | grep -v Warning | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step selftest 120 python -m pytest tests/test_experimental_gpu.py -m gpu -x -q
step gruseq   300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "gru_seq_kernel"
step ab       240 python profiles/gru_ab.py
step parity   400 python -m pytest tests/test_parity_gpu.py tests/test_chunked_gpu.py -m gpu -x -q
step tl_auto  120 python profiles/gru_timeline.py --steps 6 --impl auto
step tl_unp   120 python profiles/gru_timeline.py --steps 6 --impl unpaired
