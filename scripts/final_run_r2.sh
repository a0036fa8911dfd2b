# round 2: the command sequence behind the files under profiles/ named *_r2_* (run with gpurun on one B200)
set -x
mkdir -p gpurun_out
date; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2_final.log 2>&1; tail -4 gpurun_out/pytest_r2_final.log
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 400 gpurun_out/bench_r2_final.json
python bench.py --impl reference > gpurun_out/bench_r2_ref_final.json 2> gpurun_out/bench_r2_ref_final.err
python tests/tools/parity_campaign.py > gpurun_out/parity_campaign_r2.json 2> gpurun_out/parity_campaign_r2.err; tail -c 600 gpurun_out/parity_campaign_r2.json
python tests/tools/visibility_campaign.py > gpurun_out/visibility_campaign_r2.json 2> gpurun_out/visibility_campaign_r2.err; tail -c 400 gpurun_out/visibility_campaign_r2.json
python scripts/bench_visibility.py 10000 > gpurun_out/cvis_r2.json 2>/dev/null; python scripts/bench_visibility.py 10000 ring > gpurun_out/cvis_ring_r2.json 2>/dev/null; cat gpurun_out/cvis_r2.json gpurun_out/cvis_ring_r2.json
python scripts/profile_cycle_host.py 5 > gpurun_out/cycle_host_r2.txt 2>&1; head -2 gpurun_out/cycle_host_r2.txt; date
# launch lists (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_cycle.csv python scripts/profile_cycle.py > gpurun_out/c_ncu.log 2>&1
# full captures of every kernel of the path
ncu --set full --clock-control none --import-source on -k regex:fo_metric_sweep -s 1 -c 1 -f -o gpurun_out/prof_r2_sweep python scripts/profile_metric.py 1000000 256 51 2 > gpurun_out/p1.log 2>&1; tail -1 gpurun_out/p1.log
FO_EXACT_DCE=1 ncu --set full --clock-control none --import-source on -k regex:fo_metric_sweep -s 1 -c 1 -f -o gpurun_out/prof_r2_sweep_ties python scripts/profile_metric.py 200000 256 51 2 > gpurun_out/p1b.log 2>&1; tail -1 gpurun_out/p1b.log
ncu --set full --clock-control none --import-source on -k regex:fo_metric_detail -s 1 -c 1 -f -o gpurun_out/prof_r2_detail python scripts/bench_detail.py > gpurun_out/p2.log 2>&1; tail -1 gpurun_out/p2.log
ncu --set full --clock-control none --import-source on -k regex:fo_visibility_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_visibility python scripts/bench_visibility.py 10000 ring > gpurun_out/p3.log 2>&1; tail -1 gpurun_out/p3.log
ncu --set full --clock-control none --import-source on -k regex:fo_points_kernel -s 6 -c 2 -f -o gpurun_out/prof_r2_points python scripts/profile_cycle.py > gpurun_out/p4.log 2>&1; tail -1 gpurun_out/p4.log
ncu --set full --clock-control none --import-source on -k regex:fo_rollout_path -s 1 -c 1 -f -o gpurun_out/prof_r2_rollout python scripts/profile_cycle.py > gpurun_out/p5.log 2>&1; tail -1 gpurun_out/p5.log
compute-sanitizer --tool memcheck python scripts/sanitize_run.py > gpurun_out/sanitizer_r2_memcheck.txt 2>&1; tail -2 gpurun_out/sanitizer_r2_memcheck.txt
compute-sanitizer --tool racecheck python scripts/sanitize_run.py > gpurun_out/sanitizer_r2_racecheck.txt 2>&1; tail -2 gpurun_out/sanitizer_r2_racecheck.txt
date
