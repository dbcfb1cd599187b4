#!/bin/bash
# One GPU-box pass for the profiles/ artefacts of a round: per-launch table, ncu launch list (time + DRAM bytes) of the
# bench command, one `--set full` capture of the up_3 launches, sanitizer passes on the kernels new this round.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 300 python tools/gpu_quick_bench.py --table > gpurun_out/${TAG}_launch_table.txt 2>&1; echo "table rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches_time_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity \
  --no-reference-gpu --no-extra-configs > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 85 -c 7 -f -o gpurun_out/${TAG}_up3 \
  python tools/gpu_quick_bench.py --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${TAG}_up3.ncu-rep
timeout 600 compute-sanitizer --tool racecheck python tools/gpu_racecheck_poisson.py 6 > gpurun_out/${TAG}_racecheck_poisson.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_racecheck_poisson.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_bisenet_gpu.py tests/test_pil_resize.py -m gpu -q -x > gpurun_out/${TAG}_memcheck_bisenet.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck_bisenet.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_ct_train.py -m gpu -q -x -k "persistent or gradients" > gpurun_out/${TAG}_memcheck_cttrain.txt 2>&1; echo "memcheck ct rc=$?"; tail -4 gpurun_out/${TAG}_memcheck_cttrain.txt
