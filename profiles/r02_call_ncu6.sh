set -u
mkdir -p gpurun_out
out=gpurun_out/r02I
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 6 "${out}_${name}.log" | grep -v Warning | cut -c1-200 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step warm 100 python profiles/gru_ab.py --n 151552 --h 256 --d-in 256 --impls wide --iters 1
step ncu_wide 300 ncu --set full --clock-control none --import-source on -k regex:gru_wide_step_kernel -s 65 -c 2 -f -o gpurun_out/r02_prof_wide256 python profiles/gru_ab.py --n 151552 --h 256 --d-in 256 --impls wide --iters 1
step ncu_lin 300 ncu --set full --clock-control none --import-source on -k regex:"linear_gen_kernel|linear_tc_kernel" -s 8 -c 8 -f -o gpurun_out/r02_prof_linear python profiles/linear_ab.py
step launches_cfg5s 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_cfg5s.csv python bench.py --config cfg5s --steps 1 --warmup 3 --no-cpu-baseline
