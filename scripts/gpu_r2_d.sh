set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2d_pytest.log 2>&1; tail -15 gpurun_out/r2d_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stages > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print('sweep ms %.2f evals/s %.4g e2e %.4g lat graph %.1f clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['latency']['graph_p50_us'], d['clocks']))"
FO_EXACT_DCE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stages --no-latency > gpurun_out/r2d_bench_exact.json 2> gpurun_out/r2d_bench_exact.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_exact.json')); print('EXACT sweep ms %.2f evals/s %.4g e2e %.4g clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']))"
