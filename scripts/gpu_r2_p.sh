set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fo_metric_sweep -s 1 -c 1 -f -o gpurun_out/prof_r2p_sweep python scripts/profile_metric.py 400000 256 51 2 > gpurun_out/p1.log 2>&1; tail -1 gpurun_out/p1.log
ls -la gpurun_out/*.ncu-rep
