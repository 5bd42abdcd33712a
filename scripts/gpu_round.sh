#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms, all workloads), step profiles, ncu launch list + full captures.
# Usage (from the container): gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 rc=$?"; cat $O/${TAG}_bench_c2.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "ref rc=$?"
cat $O/${TAG}_bench_reference.json
for w in c3 c4 c4p c5 c5cs; do
  python bench.py --workload $w --steps 2 --warmup 3 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; echo "bench $w rc=$?"
  cat $O/${TAG}_bench_$w.json; tail -2 $O/${TAG}_bench_$w.err
done
python scripts/step_profile.py c2 256 bf16 > $O/${TAG}_step_profile_c2_b256.log 2>&1
python scripts/step_profile.py adm256 32 bf16 > $O/${TAG}_step_profile_adm256_b32.log 2>&1
head -4 $O/${TAG}_step_profile_c2_b256.log $O/${TAG}_step_profile_adm256_b32.log
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2100 -c 2200 --csv \
    --log-file $O/${TAG}_launches_c2_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > $O/${TAG}_ncu_run.log 2>&1; echo "ncu launch list rc=$?"
if [ "${NCU_FULL:-0}" = "1" ]; then
for k in attn_fused_kernel:0:3 gn_apply_kernel:8:3 conv_tc_kernel:30:6; do
  name=${k%%:*}; rest=${k#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  ncu --set full --clock-control none --import-source on -k regex:$name --launch-skip $skip --launch-count $cnt \
      -f -o $O/${TAG}_ncu_adm_$name python scripts/step_profile.py adm256 32 bf16 1 > $O/${TAG}_ncu_adm_$name.log 2>&1
  echo "ncu full $name rc=$?"
done
fi
