set -u
mkdir -p gpurun_out
out=gpurun_out/r02q
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 12 "${out}_${name}.log" | grep -v Warning | cut -c1-160 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step warm 100 python -c "import torch; torch.zeros(1).cuda(); print(1)"
step g1 70 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gru2_kernel -c 2 python profiles/profile_kernels.py --config cfg2
step g2 70 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gru2_kernel -c 2 python profiles/profile_kernels.py --config cfg2
step g3 70 ncu --metrics smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct --clock-control none -k regex:gru2_kernel -c 2 python profiles/profile_kernels.py --config cfg2
