#!/bin/bash
# A/B on one box: bench.py (resident value, e2e, stream block) with the product library under a few settings and with
# another build of libmadicp_b200.so dropped into scripts/_bin/<name>/.  Usage: bash scripts/ab_stream.sh <name> [tag]
NAME=${1:-r02}; TAG=${2:-ab}
LIB=mad_icp_b200/lib/libmadicp_b200.so
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" MADICP_PIPELINE_TIMING=1 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --stream-scans 1000 --stream-cpu-scans 20 > gpurun_out/${TAG}_$tag.json 2> gpurun_out/${TAG}_$tag.err
  python - <<PY
import json
l = json.loads(open("gpurun_out/${TAG}_$tag.json").read().strip().splitlines()[-1])
print("$tag: value %.1f scans/s (%.1f us)  e2e %.1f  stream %.1f scans/s (%.3f ms/scan) ate %.1e kf_equal %s" % (l["value"], 1e3 * l["ms_per_step"], l["e2e"]["value"], l["stream"]["value"], l["stream"]["ms_per_scan"], l["stream"].get("ate_m", -1), l["stream"].get("keyframe_decisions_equal")))
PY
  grep "Pipeline phases" gpurun_out/${TAG}_$tag.err; }
cp $LIB /tmp/product.so
run new1 A=1
run new_ht64 MADICP_HOST_THREADS=64
run new_la64 MADICP_LOOKAHEAD=64 MADICP_BENCH_LOOKAHEAD=64
if [ -f scripts/_bin/$NAME/libmadicp_b200.so ]; then cp scripts/_bin/$NAME/libmadicp_b200.so $LIB; run ${NAME}_1 A=1; cp /tmp/product.so $LIB; fi
run new2 A=1
