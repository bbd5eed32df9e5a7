#!/bin/bash
# one gpurun call: stage list as arguments (new | bench | ncu | full | smoke); logs under gpurun_out/
mkdir -p gpurun_out
for stage in "$@"; do
  case $stage in
    new)   timeout 420 python -m pytest tests/test_lsap.py tests/test_max_iou_assigner.py tests/test_gpu_p2p.py -q -m gpu -s > gpurun_out/new_tests.log 2>&1; echo "new rc=$?" >> gpurun_out/stages.log;;
    rpn)   timeout 300 python -m pytest tests/test_rpn.py -q -m gpu -s > gpurun_out/rpn_tests.log 2>&1; echo "rpn rc=$?" >> gpurun_out/stages.log;;
    lsap)  timeout 300 python -m pytest tests/test_lsap.py -q -m gpu -s > gpurun_out/lsap_tests.log 2>&1; echo "lsap rc=$?" >> gpurun_out/stages.log;;
    ncufull) timeout 400 ncu --set full --clock-control none --import-source on -k regex:'hungarian_v2_kernel|rpn_|miou_|lsap_prep|p2p_select' -o gpurun_out/new_kernels -f python tools/profile_new_kernels.py > gpurun_out/ncufull.log 2>&1; echo "ncufull rc=$?" >> gpurun_out/stages.log;;
    smoke) timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/stages.log;;
    bench) timeout 420 python bench.py --steps 50 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/stages.log;;
    full)  timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "full rc=$?" >> gpurun_out/stages.log;;
    ncu)   timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --profile --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/stages.log;;
  esac
done
tail -3 gpurun_out/stages.log
