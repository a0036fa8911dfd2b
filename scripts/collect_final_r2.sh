# after scripts/final_run_r2.sh: summaries of the captures and copies of the evidence into profiles/ (run in the build container)
set -e
for n in sweep sweep_ties detail visibility points rollout; do python scripts/ncu_summary.py gpurun_out/prof_r2_$n.ncu-rep > profiles/ncu_r2_${n}_summary.txt 2>/dev/null; done
for n in sweep sweep_ties; do echo "# instruction share per code region (ncu source page, scripts/sweep_regions.sh; inlined code is attributed to every line of its inline chain, so shares are relative)" >> profiles/ncu_r2_${n}_summary.txt; bash scripts/sweep_regions.sh gpurun_out/prof_r2_$n.ncu-rep >> profiles/ncu_r2_${n}_summary.txt 2>/dev/null; done
echo "# hottest source lines (scripts/ncu_lines.py)" >> profiles/ncu_r2_sweep_summary.txt; python scripts/ncu_lines.py gpurun_out/prof_r2_sweep.ncu-rep 30 2>/dev/null | cut -c1-170 >> profiles/ncu_r2_sweep_summary.txt
echo "# hottest source lines (scripts/ncu_lines.py)" >> profiles/ncu_r2_detail_summary.txt; python scripts/ncu_lines.py gpurun_out/prof_r2_detail.ncu-rep 30 2>/dev/null | cut -c1-170 >> profiles/ncu_r2_detail_summary.txt
cp gpurun_out/bench_r2_final.json gpurun_out/bench_r2_ref_final.json gpurun_out/parity_campaign_r2.json gpurun_out/visibility_campaign_r2.json gpurun_out/cvis_r2.json gpurun_out/cvis_ring_r2.json gpurun_out/cycle_host_r2.txt gpurun_out/launches_r2_final.csv gpurun_out/launches_r2_cycle.csv profiles/
cat gpurun_out/sanitizer_r2_memcheck.txt gpurun_out/sanitizer_r2_racecheck.txt > profiles/sanitizer_r2.txt
python - <<'PY'
import csv, io, json, subprocess
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_r2_sweep.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
def get(k):
    x = float(v[h.index(k)]); unit = u[h.index(k)]
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(unit, 1.0)
rd, wr, ms = get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")
json.dump({"source": "ncu --set full --clock-control none of ONE 1,000,000-trajectory launch of fo_metric_sweep_kernel<127,0,1,0> (scripts/final_run_r2.sh: scripts/profile_metric.py 1000000 256 51 2, second launch; profiles/ncu_r2_sweep_summary.txt)",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr, "algorithmic_bytes": 1065417792, "duration_ms": ms},
          open("profiles/traffic_r2.json", "w"), indent=1)
print("traffic", rd + wr, "ms", ms)
PY
