#!/bin/bash
# One GPU-box pass: full gpu tests, smoke, bench (both arms), secondary paths, ncu launch list (time + DRAM bytes).
# Writes gpurun_out/<tag>_*.
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_ref.json
timeout 600 python tools/bench_paths.py --what zencoder,shape,ct,pipeline,backend,blend,train > gpurun_out/${TAG}_paths.jsonl 2> gpurun_out/${TAG}_paths.err; echo "paths rc=$?"
timeout 300 python tools/bench_paths.py --what gen512 --B512 ${B512:-64} --steps 6 >> gpurun_out/${TAG}_paths.jsonl 2>> gpurun_out/${TAG}_paths.err; echo "gen512 rc=$?"; cat gpurun_out/${TAG}_paths.jsonl
timeout 120 python tests/_cpu_baselines.py --what blend --n 2 >> gpurun_out/${TAG}_paths.jsonl 2>> gpurun_out/${TAG}_paths.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_time_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
