set -u
mkdir -p gpurun_out
out=gpurun_out/r02n
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 4 "${out}_${name}.log" | grep -v Warning | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step warm 200 python -c "import torch; torch.zeros(1).cuda(); print(1)"
step ncu_full 500 ncu --set full --clock-control none --import-source on -k regex:"gru2_kernel|cumspmm_vec|linear_tc_kernel" -s 4 -c 4 -f -o gpurun_out/r02_prof_cfg4 python profiles/profile_kernels.py --config cfg4
step launches 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
