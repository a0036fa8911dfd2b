#!/bin/bash
# A/B of compile-time variants of the summary kernel in ONE gpurun call.
#   here:        bash scripts/ab_variants.sh build "-DFO_SW_MINB=3" "-DFO_SW_MINB=5"     # -> frenetix_occlusion_b200/libfo_var{1,2}.so
#   on the box:  gpurun -- 'bash scripts/ab_variants.sh run 2'                            # bench line per variant (+ the default build)
#   afterwards:  bash scripts/ab_variants.sh clean
# Only fo_metric_sweep.cu is recompiled; the other objects of the default build are linked as they are.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/frenetix_occlusion_b200/csrc
case "$1" in
  build)
    shift; i=0
    make -C "$CSRC" >/dev/null
    for flags in "$@"; do
      i=$((i + 1))
      nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 -I"$ROOT/include" -I"$CSRC" \
           --expt-relaxed-constexpr $flags -c "$CSRC/fo_metric_sweep.cu" -o /tmp/fo_sweep_var$i.o
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/frenetix_occlusion_b200/libfo_var$i.so" \
           "$CSRC"/fo_metric.o "$CSRC"/fo_metric_detail.o /tmp/fo_sweep_var$i.o "$CSRC"/fo_visibility.o "$CSRC"/fo_points.o \
           "$CSRC"/fo_rollout.o "$CSRC"/fo_capi.o -lcudart
      echo "variant $i: $flags"
    done ;;
  run)
    n=$2; mkdir -p "$ROOT/gpurun_out"
    for i in 0 $(seq 1 "$n"); do
      lib=$ROOT/frenetix_occlusion_b200/libfo_var$i.so
      [ "$i" = 0 ] && lib=$ROOT/frenetix_occlusion_b200/libfo_b200.so
      FO_LIB_PATH=$lib timeout 200 python "$ROOT/bench.py" --steps 5 --warmup 3 --no-cpu-baseline --no-stages \
        > "$ROOT/gpurun_out/bench_var$i.json" 2> "$ROOT/gpurun_out/bench_var$i.err" || true
      python - "$ROOT/gpurun_out/bench_var$i.json" "$i" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("variant", sys.argv[2], "evals/s %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "graph p50 us %.1f" % d["latency"]["graph_p50_us"])
except Exception as e:
    print("variant", sys.argv[2], "failed:", e)
PY
    done ;;
  clean)
    rm -f "$ROOT"/frenetix_occlusion_b200/libfo_var*.so /tmp/fo_sweep_var*.o ;;
  *) echo "usage: $0 build FLAGS... | run N | clean"; exit 2 ;;
esac
