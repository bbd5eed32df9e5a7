#!/bin/bash
# One gpurun call = a list of bounded stages (arguments); every stage logs its wall-clock seconds to gpurun_out/stages.log so a slow
# stage is visible afterwards (a cold box pages in torch / cuDNN / scipy on first use: minutes).
#   new | rpn | lsap | mirrors | smoke | bench | bench2 (2 GPUs) | full | ncu (launch list) | ncufull (ncu --set full of the session-2 kernels)
mkdir -p gpurun_out
run() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout $t "$@"
  local rc=$?
  echo "$name rc=$rc secs=$(( $(date +%s) - t0 ))" >> gpurun_out/stages.log
}
for stage in "$@"; do
  case $stage in
    new)     run new 420 bash -c 'python -m pytest tests/test_lsap.py tests/test_max_iou_assigner.py tests/test_gpu_p2p.py -q -m gpu -s --durations=8 > gpurun_out/new_tests.log 2>&1';;
    rpn)     run rpn 300 bash -c 'python -m pytest tests/test_rpn.py -q -m gpu -s --durations=8 > gpurun_out/rpn_tests.log 2>&1';;
    lsap)    run lsap 300 bash -c 'python -m pytest tests/test_lsap.py -q -m gpu -s --durations=8 > gpurun_out/lsap_tests.log 2>&1';;
    mirrors) run mirrors 200 bash -c 'python -m pytest tests/test_zz_gpu_plugin_mirrors.py -q -m gpu -s > gpurun_out/mirror_tests.log 2>&1';;
    smoke)   run smoke 200 bash -c 'python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1';;
    bench)   run bench 420 bash -c 'python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err';;
    bench2)  run bench2 420 bash -c 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err';;
    full)    run full 600 bash -c 'python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/full_tests.log 2>&1';;
    ncu)     run ncu 300 bash -c 'ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --profile --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1';;
    ncufull) run ncufull 400 bash -c "ncu --set full --clock-control none --import-source on -k regex:'hungarian_v2_kernel|rpn_|miou_|lsap_prep|p2p_select' -o gpurun_out/new_kernels -f python tools/profile_new_kernels.py > gpurun_out/ncufull.log 2>&1";;
    fullshape) run fullshape 900 bash -c 'python -m pytest tests/test_gpu_full_shape.py -q -m gpu -s --durations=12 > gpurun_out/fullshape_tests.log 2>&1';;
    memcheck) run memcheck 900 bash -c 'compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_cpr_stage.py tests/test_gpu_cpr_head.py tests/test_gpu_conv_tc.py tests/test_gpu_kernels_misc.py -q -m gpu -k "lite or mil or gfocal or general_tc or f16x2 or 8-16-32" -x > gpurun_out/sanitizer_memcheck.log 2>&1';;
    racecheck) run racecheck 900 bash -c 'compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_cpr_stage.py tests/test_gpu_cpr_head.py tests/test_gpu_kernels_misc.py -q -m gpu -k "lite or mil or gfocal" -x > gpurun_out/sanitizer_racecheck.log 2>&1';;
    refine) run refine 300 bash -c 'python tools/profile_refine.py > gpurun_out/refine_times.json 2> gpurun_out/refine.err; ncu --set full --clock-control none --import-source on -k regex:refine_fused -c 2 -o gpurun_out/refine -f python tools/profile_refine.py ncu > gpurun_out/refine_ncu.log 2>&1';;
    refinemc) run refinemc 300 bash -c 'compute-sanitizer --tool memcheck --print-limit 10 python tools/profile_refine.py ncu > gpurun_out/refine_memcheck.log 2>&1';;
    trainlist) run trainlist 400 bash -c 'ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches.csv python tools/profile_train.py 2 > gpurun_out/train_ncu.log 2>&1';;
    losstests) run losstests 600 bash -c 'python -m pytest tests/test_gpu_cpr_head.py tests/test_gpu_tower_bwd.py tests/test_grid_bags.py -q -m gpu -x -s --durations=8 > gpurun_out/loss_tests.log 2>&1';;
    lossncu) run lossncu 400 bash -c "ncu --set full --clock-control none --import-source on -k regex:'cpr_loss_bwd_tile|mil_fwd_kernel' -c 2 -o gpurun_out/lossbwd -f python tools/profile_train.py 1 > gpurun_out/lossncu.log 2>&1";;
    convtests) run convtests 400 bash -c 'python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_tower_bwd.py -q -m gpu -x -s --durations=6 > gpurun_out/conv_tests.log 2>&1';;
    benchq) run benchq 300 bash -c 'python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err';;
    conv2) run conv2 300 bash -c 'python tools/profile_conv2.py > gpurun_out/conv2_times.json 2> gpurun_out/conv2.err; PTB_CONV_CLUSTER=3 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 2 -o gpurun_out/conv_pair -f python tools/profile_conv2.py ncu > gpurun_out/conv2_ncu.log 2>&1';;
    gather2) run gather2 300 bash -c 'python tools/profile_gather2.py > gpurun_out/gather2_times.json 2> gpurun_out/gather2.err; ncu --set full --clock-control none --import-source on -k regex:bag_gather -c 2 -o gpurun_out/gather_tma -f python tools/profile_gather2.py ncu > gpurun_out/gather2_ncu.log 2>&1';;
    wgrad) run wgrad 300 bash -c "python tools/profile_wgrad.py > gpurun_out/wgrad_times.json 2> gpurun_out/wgrad.err; ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -c 2 -o gpurun_out/wgrad_pair -f python tools/profile_wgrad.py ncu > gpurun_out/wgrad_ncu.log 2>&1";;
    rpnprof) run rpnprof 300 bash -c "python tools/profile_rpn.py > gpurun_out/rpn_times.json 2> gpurun_out/rpn_prof.err; ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/rpn_launches.csv python tools/profile_rpn.py ncu > gpurun_out/rpn_ncu.log 2>&1";;
    lsapprof) run lsapprof 300 bash -c "python tools/profile_lsap.py > gpurun_out/lsap_times.json 2> gpurun_out/lsap_prof.err";;
    memcheck2) run memcheck2 900 bash -c "compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_rpn.py tests/test_gpu_tower_bwd.py -q -m gpu -x > gpurun_out/sanitizer_memcheck_r2b.log 2>&1; compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_rpn.py -q -m gpu -x -k \"bitmask or kernel\" > gpurun_out/sanitizer_racecheck_rpn.log 2>&1";;
    conv3) run conv3 300 bash -c "python tools/profile_conv2.py > gpurun_out/conv2_times.json 2> gpurun_out/conv2.err";;
    stagetests) run stagetests 400 bash -c 'python -m pytest tests/test_gpu_cpr_stage.py tests/test_gpu_kernels_misc.py tests/test_grid_bags.py -q -m gpu -x -s > gpurun_out/stage_tests.log 2>&1';;
    *) echo "unknown stage $stage" >> gpurun_out/stages.log;;
  esac
done
tail -5 gpurun_out/stages.log
