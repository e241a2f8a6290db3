# Parity tests + bench (optionally for several scatter spreads) + ncu captures.  Run under gpurun.
#   bash scripts/ab_bench.sh <tag> "<spread> ..." "<kernel> ..."
mkdir -p gpurun_out
TAG=${1:-ab}
SPREADS=${2:-"8"}
KERNELS=${3:-""}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
for v in $SPREADS; do
  timeout 600 python bench.py --spread $v --no-cpu-baseline --no-downstream --e2e-steps 1 > gpurun_out/${TAG}_bench_s$v.json 2> gpurun_out/${TAG}_bench_s$v.err
  echo "spread $v bench exit $?"; tail -3 gpurun_out/${TAG}_bench_s$v.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_s$v.json'));print(d['ms_per_step'],d['kernel_ms'],d['roofline']['frac'])"
done
for K in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s 3 -c 1 -f \
      -o gpurun_out/${TAG}_$K python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-downstream > gpurun_out/${TAG}_ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
