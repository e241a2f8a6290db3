# A/B of the K2 variants (HG_OPT_K2_VARIANT): parity tests, bench, ncu capture.  Run under gpurun.
mkdir -p gpurun_out
TAG=${1:-ab}
VARIANTS=${2:-"0 1 2"}
HINGE_B200_K2_VARIANT=0 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_v0.log 2>&1
echo "variant 0 pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_v0.log
for v in $VARIANTS; do
  timeout 600 python bench.py --k2-variant $v --no-cpu-baseline --no-downstream --e2e-steps 1 > gpurun_out/${TAG}_bench_v$v.json 2> gpurun_out/${TAG}_bench_v$v.err
  echo "variant $v bench exit $?"
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_v$v.json'));print(d['ms_per_step'],d['kernel_ms'])"
done
for v in ${3:-0}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_mask_anno_flat$" -s 3 -c 1 -f \
      -o gpurun_out/${TAG}_k_mask_anno_flat_v$v python bench.py --k2-variant $v --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-downstream > gpurun_out/${TAG}_ncu_v$v.log 2>&1
  echo "ncu v$v exit $?"
done
