#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list, ncu --set full of the hot kernels.
# Usage (from the repo root, under gpurun):  bash scripts/profile_gpu.sh [tag]
# Everything lands in gpurun_out/<tag>_*; scripts/summarise_profiles.py turns it into profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log

echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log

echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json

echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "reference exit $?"
cat $OUT/${TAG}_bench_reference.json

BENCH_SHORT="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-downstream"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH_SHORT > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"

echo "== ncu --set full (hot kernels, steady-state launches)"
for K in k_profile_flat k_mask_anno_flat k_hinge_call k_hinge_exact; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s 3 -c 1 -f \
        -o $OUT/${TAG}_$K $BENCH_SHORT > $OUT/${TAG}_ncu_$K.log 2>&1
    echo "ncu $K exit $?"
done
ls -la $OUT
