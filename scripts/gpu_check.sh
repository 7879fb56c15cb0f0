#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch list + full capture of the dominant kernel.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r03}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
MADICP_PIPELINE_TIMING=1 timeout 900 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 6000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "${REF:-1}" = "1" ]; then
timeout 300 python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_reference.json
fi
timeout 300 python scripts/memo_probe.py > gpurun_out/${TAG}_memo_probe.txt 2>&1
timeout 300 python scripts/tail_probe.py > gpurun_out/${TAG}_tail_probe.txt 2>&1
MADICP_BUILD_TIMING=1 timeout 200 python scripts/build_probe_gpu.py > gpurun_out/${TAG}_build_probe.txt 2>&1
MADICP_BUILD_TIMING=1 timeout 200 python scripts/batch_probe.py 1 4 16 32 > gpurun_out/${TAG}_batch_probe.txt 2>&1
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --stream-scans 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
if [ "${NCU_FOREST:-1}" = "1" ]; then
PROBE_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_forest_launches.csv \
    python scripts/batch_probe.py 32 > gpurun_out/${TAG}_ncu_forest.log 2>&1
fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gn_loop -s 3 -c 2 -f -o gpurun_out/${TAG}_gn_loop \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --stream-scans 0 > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
ls -la gpurun_out | tail -20
