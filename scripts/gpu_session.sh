#!/bin/bash
# One GPU-box session of round 2 (run under gpurun from the repo root):
#   bash scripts/gpu_session.sh <tag> [steps...]
# steps: pytest smoke bench bench_old launches ncu_k1 ncu_all sanitize
set -u
TAG=${1:-r02}
shift || true
STEPS=${*:-"pytest smoke bench"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; free -g | head -2 >> $OUT/${TAG}_gpu.txt; df -h /tmp | tail -1 >> $OUT/${TAG}_gpu.txt
SHORT="--steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-downstream --no-verify"
for S in $STEPS; do
  case $S in
    pytest)
      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
      echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
      echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log ;;
    bench)
      timeout 1200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
      echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err; tail -c 2500 $OUT/${TAG}_bench.json ;;
    bench_ref)
      timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
      echo "reference exit $?"; cat $OUT/${TAG}_bench_reference.json ;;
    bench_old)
      timeout 900 python bench.py --profile-kernel 1 --no-cpu-baseline --no-downstream --no-verify --e2e-steps 1 > $OUT/${TAG}_bench_k1old.json 2> $OUT/${TAG}_bench_k1old.err
      echo "bench (old K1) exit $?"
      python -c "
import json;d=json.load(open('$OUT/${TAG}_bench_k1old.json'));print('c5',d['ms_per_step'],d['kernel_ms'],d['roofline']['frac']);print('c3',d['c3']['ms_per_step'],d['c3']['kernel_ms'],d['c3']['roofline']['frac'])" ;;
    pytest_filter)
      timeout 900 python -m pytest tests/test_gpu_filter_api.py tests/test_gpu_edge_cases.py tests/test_gpu_filter_cli.py -m gpu -x -q > $OUT/${TAG}_pytest_filter.log 2>&1
      echo "pytest (filter) exit $?" | tee -a $OUT/${TAG}_pytest_filter.log; tail -5 $OUT/${TAG}_pytest_filter.log ;;
    variants)
      for V in ${VARIANTS:-1 2 0}; do
        timeout 600 python bench.py --profile-kernel $V $SHORT > $OUT/${TAG}_bench_v$V.json 2> $OUT/${TAG}_bench_v$V.err
        echo "variant $V exit $?"
        python -c "
import json;d=json.load(open('$OUT/${TAG}_bench_v$V.json'));print('c5',round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['kernel_ms'].items()},round(d['roofline']['frac'],3));print('c3',round(d['c3']['ms_per_step'],4),{k:round(v,4) for k,v in d['c3']['kernel_ms'].items()},round(d['c3']['roofline']['frac'],3))"
      done ;;
    spreads)
      for S in ${SPREADS:-1 4 16}; do
        timeout 600 python bench.py --config c5 --also "" --spread $S --profile-kernel 1 $SHORT > $OUT/${TAG}_bench_s$S.json 2> $OUT/${TAG}_bench_s$S.err
        python -c "
import json;d=json.load(open('$OUT/${TAG}_bench_s$S.json'));print('spread $S c5',round(d['ms_per_step'],4),{k:round(v,4) for k,v in d['kernel_ms'].items()})"
      done ;;
    down)
      for CFG in c3 c5; do
        HINGE_B200_TIMING=1 timeout 900 python bench.py --config $CFG --also "" --steps 3 --e2e-steps 1 --no-cpu-baseline --no-verify > $OUT/${TAG}_down_$CFG.json 2> $OUT/${TAG}_down_$CFG.err
        echo "downstream $CFG exit $?"; grep "timing\]" $OUT/${TAG}_down_$CFG.err | tail -40
        python -c "
import json;d=json.load(open('$OUT/${TAG}_down_$CFG.json'));print(d.get('downstream_stages'))"
      done ;;
    pytest_pipeline)
      timeout 900 python -m pytest tests/test_gpu_pipeline_cli.py tests/test_gpu_edge_cases.py -m gpu -x -q > $OUT/${TAG}_pytest_pipeline.log 2>&1
      echo "pytest (pipeline) exit $?" | tee -a $OUT/${TAG}_pytest_pipeline.log; tail -15 $OUT/${TAG}_pytest_pipeline.log ;;
    launches_down)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file $OUT/${TAG}_launches_down.csv python bench.py --config c3 --also "" --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-verify > $OUT/${TAG}_launches_down.log 2>&1
      echo "launch list (downstream) exit $?" ;;
    ncu_down)
      for K in ${NCU_DOWN:-k_classify_reads k_layout_pairs k_order_candidates k_hinge_exact_warp}; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K" -s 1 -c 1 -f \
          -o $OUT/${TAG}_${K}_c3 python bench.py --config c3 --also "" --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-verify > $OUT/${TAG}_ncu_${K}.log 2>&1
        echo "ncu $K exit $?"
      done ;;
    sharded)
      for SHAPE in c3 c5; do
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 \
          scripts/sharded_check.py $SHAPE > $OUT/${TAG}_sharded_$SHAPE.log 2>&1
        echo "sharded_check $SHAPE exit $?"; grep SHARDED_CHECK $OUT/${TAG}_sharded_$SHAPE.log; tail -3 $OUT/${TAG}_sharded_$SHAPE.log
      done ;;
    bench2)
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 \
        bench.py --gpus 2 --no-downstream > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
      echo "bench N=2 exit $?"; tail -5 $OUT/${TAG}_bench_n2.err; tail -c 1500 $OUT/${TAG}_bench_n2.json
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 \
        bench.py --gpus 2 --no-downstream --exchange nccl --no-verify > $OUT/${TAG}_bench_n2_nccl.json 2> $OUT/${TAG}_bench_n2_nccl.err
      echo "bench N=2 nccl exit $?"
      python -c "
import json
for f in ('$OUT/${TAG}_bench_n2.json','$OUT/${TAG}_bench_n2_nccl.json'):
    d=json.load(open(f)); print(f, 'c5', round(d['ms_per_step'],4), d['kernel_ms'], d.get('parity'), 'c3', round(d['c3']['ms_per_step'],4), d['c3'].get('parity'))" ;;
    benchN)
      NG=${NGPUS:-4}
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29651 \
        bench.py --gpus $NG --no-downstream > $OUT/${TAG}_bench_n$NG.json 2> $OUT/${TAG}_bench_n$NG.err
      echo "bench N=$NG exit $?"; grep -v "^\*\|^$\|OMP_NUM" $OUT/${TAG}_bench_n$NG.err | tail -5
      python -c "
import json
d=json.load(open('$OUT/${TAG}_bench_n$NG.json')); print('c5 value %.4g ms %.4f'%(d['value'],d['ms_per_step']), d['kernel_ms'], d.get('parity'), d['config']['overlaps_rank0']); print('c3 value %.4g ms %.4f'%(d['c3']['value'],d['c3']['ms_per_step']), d['c3']['kernel_ms'], d['c3'].get('parity')); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
        --log-file $OUT/${TAG}_launches.csv python bench.py $SHORT > $OUT/${TAG}_launches_bench.log 2>&1
      echo "launch list exit $?" ;;
    ncu_k1)
      for CFG in c5 c3; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_profile_flat2" -s 4 -c 1 -f \
          -o $OUT/${TAG}_k_profile_flat2_$CFG python bench.py --config $CFG --also "" $SHORT > $OUT/${TAG}_ncu_k1_$CFG.log 2>&1
        echo "ncu k_profile_flat2 $CFG exit $?"
      done ;;
    ncu_one)
      # NCU_K=<kernel regex> NCU_CFG=c3|c5 [NCU_ARGS="--profile-kernel 1"]
      K=${NCU_K:-k_profile_tma}; CFG=${NCU_CFG:-c3}
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K" -s 4 -c 1 -f \
        -o $OUT/${TAG}_${K}_$CFG python bench.py --config $CFG --also "" $SHORT ${NCU_ARGS:-} > $OUT/${TAG}_ncu_${K}_$CFG.log 2>&1
      echo "ncu $K $CFG exit $?" ;;
    ncu_k4)
      for K in k_hinge_exact_warp k_hinge_call; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K" -s 4 -c 1 -f \
          -o $OUT/${TAG}_${K}_c5 python bench.py --config c5 --also "" $SHORT > $OUT/${TAG}_ncu_${K}.log 2>&1
        echo "ncu $K exit $?"
      done ;;
    ncu_maximal)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_classify_reads" -s 1 -c 1 -f \
        -o $OUT/${TAG}_k_classify_reads_c3 python bench.py --config c3 --also "" --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-verify > $OUT/${TAG}_ncu_k_classify_reads.log 2>&1
      echo "ncu k_classify_reads exit $?" ;;
    ncu_k1final)
      for SPEC in "k_profile_flat2 c3" "k_profile_flat c5"; do
        set -- $SPEC; K=$1; CFG=$2
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\\b" -s 4 -c 1 -f \
          -o $OUT/${TAG}_${K}_$CFG python bench.py --config $CFG --also "" $SHORT > $OUT/${TAG}_ncu_${K}_$CFG.log 2>&1
        echo "ncu $K $CFG exit $?"
      done ;;
    ncu_final)
      # the kernels of a filter step on both workloads, one launch each (steady state: -s skips the warm-up)
      for SPEC in "k_profile_flat2 c3" "k_profile_flat c5" "k_mask_bits_flat c5" "k_mask_walk c5" "k_mask_bits_flat c3" \
                  "k_mask_walk c3" "k_hinge_call c5" "k_hinge_exact_warp c5" "k_hinge_call c3"; do
        set -- $SPEC; K=$1; CFG=$2
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\b" -s 4 -c 1 -f \
          -o $OUT/${TAG}_${K}_$CFG python bench.py --config $CFG --also "" $SHORT > $OUT/${TAG}_ncu_${K}_$CFG.log 2>&1
        echo "ncu $K $CFG exit $?"
      done
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_profile_tma" -s 4 -c 1 -f \
        -o $OUT/${TAG}_k_profile_tma_c3 python bench.py --config c3 --also "" --profile-kernel 3 $SHORT > $OUT/${TAG}_ncu_k_profile_tma_c3.log 2>&1
      echo "ncu k_profile_tma c3 exit $?"
      timeout 600 python bench.py --profile-kernel 3 $SHORT > $OUT/${TAG}_bench_tma.json 2> $OUT/${TAG}_bench_tma.err
      echo "bench (TMA form of K1) exit $?" ;;
    ncu_all)
      for K in k_mask_anno_flat k_hinge_call k_hinge_exact_warp k_profile_flat; do
        timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K" -s 4 -c 1 -f \
          -o $OUT/${TAG}_${K}_c5 python bench.py --config c5 --also "" $SHORT > $OUT/${TAG}_ncu_${K}.log 2>&1
        echo "ncu $K exit $?"
      done ;;
    sanitize2)
      # memcheck and racecheck of one small filter run (smoke) with the default forms of K1 and with the TMA-staged one
      for TOOL in memcheck racecheck; do
        for PK in 0 3; do
          HINGE_B200_PROFILE_KERNEL=$PK timeout 900 compute-sanitizer --tool $TOOL python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_${TOOL}_pk$PK.log 2>&1
          echo "$TOOL (profile kernel $PK) exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/${TAG}_${TOOL}_pk$PK.log | tail -3
        done
      done ;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_memcheck_smoke.log 2>&1
      echo "memcheck exit $?"; tail -3 $OUT/${TAG}_memcheck_smoke.log ;;
  esac
done
ls -la $OUT | tail -30
