#!/bin/bash
# Round-2 GPU session script (run under gpurun from the repo root).  Usage: tools/gpu_call.sh <tag> <stage...>
set -u
TAG=$1; shift
NG=${NG:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for stage in "$@"; do
  case $stage in
    tests)     timeout 900 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?" ;;
    newtests)  timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q > $OUT/newtests.log 2>&1; echo "newtests rc=$?" ;;
    scoreerr)  timeout 600 python tools/tc_score_error.py > $OUT/tc_score_error.txt 2> $OUT/tc_score_error.err; echo "scoreerr rc=$?" ;;
    bench)     timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" ;;
    bench_tau) for s in 0.5 0.25 0.125; do LEIDA_TC_TAU_SCALE=$s timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --no-secondary > $OUT/bench_tau$s.json 2> $OUT/bench_tau$s.err; echo "bench_tau$s rc=$?"; done ;;
    bench2)    timeout 600 python bench.py --config config2 --steps 5 --warmup 3 --no-cpu > $OUT/bench_config2.json 2> $OUT/bench_config2.err; echo "bench2 rc=$?" ;;
    ref)       timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?" ;;
    ncu_list)  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > $OUT/ncu_list.log 2>&1; echo "ncu_list rc=$?" ;;
    mgpu)      for x in auto peer nccl; do LEIDA_EXCHANGE=$x timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > $OUT/mgpu_check_${NG}gpu_$x.log 2>&1; echo "mgpu $x rc=$?"; grep "world=" $OUT/mgpu_check_${NG}gpu_$x.log; done ;;
    mbench)    for x in ${XCHG:-auto nccl}; do LEIDA_EXCHANGE=$x timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 3 --warmup 2 > $OUT/bench_${NG}gpu_$x.json 2> $OUT/bench_${NG}gpu_$x.err; echo "mbench $x rc=$?"; done ;;
    stage12)   timeout 600 python tools/stage12_probe.py > $OUT/stage12_probe.txt 2>&1; echo "stage12 rc=$?"; cat $OUT/stage12_probe.txt ;;
    bench_dump) LEIDA_BENCH_DUMP=$OUT/per_launch.json timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu --no-secondary > $OUT/bench_dump.json 2> $OUT/bench_dump.err; echo "bench_dump rc=$?" ;;
    ncu_stage12) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hilbert|k_eigvec_rank2" -c 4 -o $OUT/stage12 -f python tools/stage12_probe.py 60 > $OUT/ncu_stage12.log 2>&1; echo "ncu_stage12 rc=$?" ;;
    config4)   timeout 900 python bench.py --config config4 --steps 2 --warmup 1 --no-secondary > $OUT/bench_config4.json 2> $OUT/bench_config4.err; echo "config4 rc=$?"; timeout 600 python tools/config4_dense.py 3 > $OUT/config4_dense.txt 2>&1; echo "config4_dense rc=$?"; cat $OUT/config4_dense.txt ;;
    config5)   timeout 1200 python bench.py --config config5 --steps 1 --warmup 1 --no-secondary > $OUT/bench_config5.json 2> $OUT/bench_config5.err; echo "config5 rc=$?" ;;
    ncu_final) B="python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary --no-e2e"
               timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv $B > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
               timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_assign|k_recheck|k_diff_apply|k_update" -c 4 -o $OUT/pass0 -f $B > $OUT/ncu_pass0.log 2>&1; echo "ncu pass0 rc=$?"
               timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_distortion_tiled" -c 1 -o $OUT/distortion -f $B > $OUT/ncu_dist.log 2>&1; echo "ncu distortion rc=$?"
               timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hilbert|k_eigvec_rank2" -c 2 -o $OUT/stage12 -f $B > $OUT/ncu_s12.log 2>&1; echo "ncu stage12 rc=$?" ;;
    mgpu_auto) LEIDA_EXCHANGE=auto timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > $OUT/mgpu_check_${NG}gpu_auto.log 2>&1; echo "mgpu auto rc=$?"; grep "world=" $OUT/mgpu_check_${NG}gpu_auto.log ;;
    mbench4)   LEIDA_EXCHANGE=auto timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 2 > $OUT/bench_4gpu_auto.json 2> $OUT/bench_4gpu_auto.err; echo "mbench4 rc=$?" ;;
    mconfig5)  LEIDA_EXCHANGE=auto timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NG --config config5 --steps 1 --warmup 1 > $OUT/bench_config5_${NG}gpu.json 2> $OUT/bench_config5_${NG}gpu.err; echo "mconfig5 rc=$?" ;;
    bn256)     export LEIDA_B200_LIB=$PWD/leida-python_b200/pyleida/_lib/libleida_b200_bn256.so
               timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core_and_cuda_core or kmeans_golden or pipeline_golden" > $OUT/tests_bn256.log 2>&1; rc=$?; echo "tests bn256 rc=$rc"; tail -3 $OUT/tests_bn256.log
               if [ $rc -eq 0 ]; then
                 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary --no-e2e > $OUT/bench_bn256.json 2> $OUT/bench_bn256.err; echo "bench bn256 rc=$?"
                 timeout 100 python bench.py --config config2 --steps 5 --warmup 3 --no-cpu --no-secondary --no-e2e > $OUT/bench_bn256_config2.json 2> $OUT/bench_bn256_config2.err; echo "bench bn256 config2 rc=$?"
               fi
               unset LEIDA_B200_LIB ;;
    variant)   # A/B of a variant library (VARIANT=name) against the default build: kernel-level bench lines only
               for lib in default $VARIANT; do
                 if [ $lib = default ]; then unset LEIDA_B200_LIB; else export LEIDA_B200_LIB=$PWD/leida-python_b200/pyleida/_lib/libleida_b200_$lib.so; fi
                 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu --no-secondary --no-e2e > $OUT/bench_$lib.json 2> $OUT/bench_$lib.err; echo "bench $lib rc=$?"
               done; unset LEIDA_B200_LIB ;;
    ncu_list_final) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu --no-secondary --no-e2e > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" ;;
    quick)     timeout 60 python bench.py --steps 3 --warmup 2 --no-cpu --no-secondary --no-e2e > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench rc=$?"
               timeout 100 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/tests.log ;;
    final)     timeout 600 python -m pytest tests -m gpu -x -q > $OUT/tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/tests.log
               timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
               timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" ;;
    dist_ab)   # same-box A/B of the two ring forms of the distortion kernel, alternating
               timeout 200 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k distortion_forms > $OUT/tests_forms.log 2>&1; rc=$?; echo "forms rc=$rc"; tail -3 $OUT/tests_forms.log
               [ $rc -eq 0 ] && for i in 1 2; do for f in ${FORMS:-piped ring}; do
                 LEIDA_DISTORTION_FORM=$f timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu --no-secondary --no-e2e > $OUT/bench_${f}_$i.json 2> $OUT/bench_${f}_$i.err; echo "bench $f $i rc=$?"
               done; done ;;
    driver)    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_driver.json 2> $OUT/bench_driver.err; echo "driver-style bench rc=$?" ;;
    *) echo "unknown stage $stage" ;;
  esac
done
tail -n 5 $OUT/*.log 2>/dev/null | cut -c1-300
