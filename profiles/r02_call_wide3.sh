set -u
mkdir -p gpurun_out
out=gpurun_out/r02y
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; tail -n 14 "${out}_${name}.log" | cut -c1-300 | sed "s/^/    /" | tee -a "${out}_summary.log"; }
step tl256 120 python profiles/gru_wide_timeline.py --h 256 --d-in 256 --steps 4
step tl256_m1 120 python profiles/gru_wide_timeline.py --h 256 --d-in 256 --steps 4 --mode 1
CTGCN_WIDE_UNITS=8 step tl256_u8 120 python profiles/gru_wide_timeline.py --h 256 --d-in 256 --steps 4
step tl128 120 python profiles/gru_wide_timeline.py --h 128 --d-in 128 --steps 4
