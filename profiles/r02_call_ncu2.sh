set -u
mkdir -p gpurun_out
out=gpurun_out/r02o
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 6 "${out}_${name}.log" | grep -v Warning | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step plain   200 python profiles/profile_kernels.py --config cfg2
step ncu_triv 200 ncu --metrics gpu__time_duration.sum python -c "import torch; x=torch.zeros(1024,device='cuda'); y=x+1; torch.cuda.synchronize(); print('ok')"
step ncu_cfg2 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gru2_kernel|cumspmm_vec|linear_tc_kernel" python profiles/profile_kernels.py --config cfg2
which ncu; ncu --version | tail -2
