set -u
mkdir -p gpurun_out
out=gpurun_out/r02k
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 14 "${out}_${name}.log" | grep -v Warning | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step cd_small 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "core_diffusion_golden and auto"
step tl_cfg4  200 python profiles/cd_timeline.py --config cfg4
step ab_cfg4  300 python profiles/cd_ab.py --config cfg4
