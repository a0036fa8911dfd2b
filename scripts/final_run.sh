set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err
python bench.py --impl reference > gpurun_out/bench_r1_ref_final2.json 2> gpurun_out/bench_r1_ref_final2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fo_metric_sweep -s 1 -c 1 -f -o gpurun_out/prof_r1_final2 python scripts/profile_metric.py 1000000 256 51 2 > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
