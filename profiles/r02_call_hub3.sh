set -u
mkdir -p gpurun_out
out=gpurun_out/r02F
step() { local name=$1 limit=$2; shift 2; local t0=$SECONDS; timeout "$limit" "$@" > "${out}_${name}.log" 2>&1; local rc=$?
  echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"; grep '^{' "${out}_${name}.log" | tail -n 1 | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('   ', round(j['ms_per_step'], 2), 'ms/step', {k: (round(v['avg_ms'], 3), v['launches'], round(v['frac'], 3)) for k, v in j['roofline_by_kernel'].items()})" | tee -a "${out}_summary.log"; }
for t in 512 1024 2048 8192; do CTGCN_HUB_THRESHOLD=$t step cfg5s_t$t 300 python bench.py --config cfg5s --steps 3 --warmup 3; done
