#!/usr/bin/env bash
# ncu evidence of a round: launch list of one replayed iteration + --set full captures of the kept kernels.
#   gpurun --timeout 2400 -- 'bash tools/profile_round.sh r02'
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-step --warmup 3 --distinct 4 > gpurun_out/${tag}_launches.log 2>&1
wc -l gpurun_out/${tag}_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 4 -o gpurun_out/${tag}_conv_resblock \
    python tools/conv_probe.py --once > gpurun_out/${tag}_ncu_conv.log 2>&1; tail -1 gpurun_out/${tag}_ncu_conv.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:layout_fwd_tile -c 2 -o gpurun_out/${tag}_layout_cfg4 \
    python tools/hbm_kernels.py --config cfg4 --reps 1 > gpurun_out/${tag}_ncu_layout4.log 2>&1; tail -1 gpurun_out/${tag}_ncu_layout4.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:layout_fwd_tile -c 2 -o gpurun_out/${tag}_layout_cfg2 \
    python tools/hbm_kernels.py --config cfg2 --reps 1 > gpurun_out/${tag}_ncu_layout2.log 2>&1; tail -1 gpurun_out/${tag}_ncu_layout2.log
timeout 400 ncu --set full --clock-control none -k regex:"gather_concat|pool_kernel|pool_bwd|gather_bwd|layout_bwd_vecs|crop_" -c 12 \
    -o gpurun_out/${tag}_graph_cfg5 python tools/hbm_kernels.py --config cfg5 --reps 1 > gpurun_out/${tag}_ncu_graph.log 2>&1; tail -1 gpurun_out/${tag}_ncu_graph.log
for c in cfg2 cfg4 cfg5; do timeout 200 python tools/hbm_kernels.py --config $c --json gpurun_out/${tag}_hbm_$c.json | tail -13; done
