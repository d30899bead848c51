set -u
mkdir -p gpurun_out
out=gpurun_out/r02V
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 3 "${out}_${name}.log" | grep -v Warning | cut -c1-200 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step warm 100 python profiles/profile_kernels.py --config cfg2
step ncu_cfg4 240 ncu --set full --clock-control none --import-source on -k regex:"gru2_kernel|linear_tc_kernel" -s 3 -c 3 -f -o gpurun_out/r02_prof_cfg4_final python profiles/profile_kernels.py --config cfg4
