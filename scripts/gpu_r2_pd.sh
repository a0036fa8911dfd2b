set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fo_metric_detail -s 1 -c 1 -f -o gpurun_out/prof_r2p_detail python scripts/bench_detail.py > gpurun_out/p2.log 2>&1; tail -1 gpurun_out/p2.log
