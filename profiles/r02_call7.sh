set -u
mkdir -p gpurun_out
out=gpurun_out/r02g
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 8 "${out}_${name}.log" | grep -v Warning | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step cd_small 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "core_diffusion_golden and auto"
step ab_cfg2  200 python profiles/cd_ab.py --config cfg2
step ab_cfg4  300 python profiles/cd_ab.py --config cfg4
step parity   400 python -m pytest tests/test_parity_gpu.py tests/test_chunked_gpu.py tests/test_uci_e2e.py -m gpu -x -q
step bench    300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
