set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -k "not scenario" > gpurun_out/r2c_pytest.log 2>&1; tail -15 gpurun_out/r2c_pytest.log
bash scripts/ab_variants.sh each 2 python scripts/bench_detail.py 2>&1 | grep -v "^+" | tee gpurun_out/r2c_detail_variants.log
AB_LIST="0 5" bash scripts/ab_variants.sh each 0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-stages --no-latency 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('=='): print(line); continue
    try:
        d=json.loads(line); print('  sweep ms %.2f evals/s %.4g e2e %.4g' % (d['ms_per_step'], d['value'], d['e2e']['value']))
    except Exception: pass
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fo_metric_detail -s 1 -c 1 -f -o gpurun_out/prof_r2c_detail python scripts/bench_detail.py > gpurun_out/r2c_prof.log 2>&1; tail -2 gpurun_out/r2c_prof.log
