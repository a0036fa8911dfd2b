#!/bin/bash
# A/B of compile-time variants of ONE kernel source in ONE gpurun call.
#   here:        bash scripts/ab_variants.sh build fo_metric_sweep.cu "-DFO_SW_MINB=3" "-DFO_SW_MINB=5"
#                                                        # -> frenetix_occlusion_b200/libfo_var{1,2}.so
#   on the box:  gpurun -- 'bash scripts/ab_variants.sh run 2'            # bench.py line per variant (+ the default build)
#                gpurun -- 'bash scripts/ab_variants.sh each 2 python scripts/bench_detail.py'   # any command per variant
#   afterwards:  bash scripts/ab_variants.sh clean
# Only the named source is recompiled; the other objects of the default build are linked as they are.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/frenetix_occlusion_b200/csrc
ALL="fo_metric fo_metric_detail fo_metric_sweep fo_visibility fo_points fo_spawn fo_rollout fo_capi"
case "$1" in
  build)
    shift; src=$1; shift; i=${AB_FIRST:-1}; i=$((i - 1))     # AB_FIRST=5: number the variants from 5 on
    base=${src%.cu}
    make -C "$CSRC" >/dev/null
    for flags in "$@"; do
      i=$((i + 1))
      nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 -I"$ROOT/include" -I"$CSRC" \
           --expt-relaxed-constexpr $flags -c "$CSRC/$src" -o /tmp/fo_var$i.o
      objs=""
      for o in $ALL; do
        if [ "$o" = "$base" ]; then objs="$objs /tmp/fo_var$i.o"; else objs="$objs $CSRC/$o.o"; fi
      done
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/frenetix_occlusion_b200/libfo_var$i.so" $objs -lcudart
      echo "variant $i: $src $flags"
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
  each)
    n=$2; shift; shift
    for i in ${AB_LIST:-0 $(seq 1 "$n")}; do
      lib=$ROOT/frenetix_occlusion_b200/libfo_var$i.so
      [ "$i" = 0 ] && lib=$ROOT/frenetix_occlusion_b200/libfo_b200.so
      echo "== variant $i"
      FO_LIB_PATH=$lib timeout 300 "$@" || true
    done ;;
  clean)
    rm -f "$ROOT"/frenetix_occlusion_b200/libfo_var*.so /tmp/fo_var*.o ;;
  *) echo "usage: $0 build SRC.cu FLAGS... | run N | each N CMD... | clean"; exit 2 ;;
esac
