set -x
mkdir -p gpurun_out
compute-sanitizer --tool memcheck python scripts/sanitize_run.py > gpurun_out/sanitizer_r2_memcheck.txt 2>&1; tail -n 4 gpurun_out/sanitizer_r2_memcheck.txt
compute-sanitizer --tool racecheck python scripts/sanitize_run.py > gpurun_out/sanitizer_r2_racecheck.txt 2>&1; tail -n 4 gpurun_out/sanitizer_r2_racecheck.txt

