# ncu launch list + full captures of the named kernels over a short bench run.  Run under gpurun.
#   bash scripts/prof_kernel.sh <tag> "<kernel> <kernel> ..."
TAG=$1; shift
KERNELS=$1
OUT=gpurun_out; mkdir -p $OUT
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-downstream"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH_SHORT > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list exit $?"
for K in $KERNELS; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s 3 -c 1 -f \
        -o $OUT/${TAG}_$K $BENCH_SHORT > $OUT/${TAG}_ncu_$K.log 2>&1
    echo "ncu $K exit $?"
done
