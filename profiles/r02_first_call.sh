#!/usr/bin/env bash
# First GPU call of round 2 in ONE gpurun invocation (runbook: profiles/r02_runbook.md).  Every step has its own timeout and
# writes into gpurun_out/, so a step that traps or hangs costs its own limit, not the call.
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/r02_first_call.sh'
#
# Order: the green-suite check first (so that the round starts from a known state), then the unmeasured experimental modes at
# configs[1] size (numerics + first timings), then configs[3] size, then the bench line and the launch list of the default path.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02a
step() {   # step <name> <seconds> <command...>
    local name=$1 limit=$2
    shift 2
    local t0=$SECONDS
    timeout "$limit" "$@" > "${out}_${name}.log" 2>&1
    local rc=$?
    echo "[$name] rc=$rc $((SECONDS - t0))s" | tee -a "${out}_summary.log"
    tail -n 6 "${out}_${name}.log" | sed "s/^/    /" | tee -a "${out}_summary.log"
}

step tests      240 python -m pytest tests -m gpu -x -q
# kernels written after the round-1 GPU budget was spent (negative-sampling loss; the 2-CTA MMA building block):
# opt-in until green once.  Separate processes: a trap in one must not take the other down.
step loss_tests 120 env CTGCN_UNVERIFIED_GPU_TESTS=1 python -m pytest tests/test_loss_gpu.py -m gpu -q
step pair_umma   90 env CTGCN_UNVERIFIED_GPU_TESTS=1 python -m pytest tests/test_experimental_gpu.py -m gpu -q
# every experimental core-GRU build in a process of its own (a trap in one must not cost the others); checksums of modes 3, 5, 6 must agree
for m in 2 3 5 6; do
    step "mode${m}_cfg2" 60 python profiles/try_mode.py --mode "$m" --config cfg2
done
for m in 2 3 5 6; do
    step "mode${m}_cfg4" 90 python profiles/try_mode.py --mode "$m" --config cfg4 --iters 5
done
step coop_cfg2  150 python profiles/try_coop.py --config cfg2
step coop_cfg4  240 python profiles/try_coop.py --config cfg4 --iters 5
step hubsplit   150 python profiles/try_hubsplit.py
# context number: the reference's own library calls (cuSPARSE / cuDNN through torch) on the GPU next to this repo's forward
step library    150 python profiles/library_gpu_baseline.py --config cfg2
step bench_cfg4 200 python bench.py --steps 10 --warmup 3
# launch list of the default bench command's timed region (per-launch times under ncu are cold-cache: compare SHARES)
step launches   200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file "${out}_launches_cfg4.csv" python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat "${out}_summary.log"
