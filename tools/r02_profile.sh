#!/bin/bash
# One GPU call: parity tests, layer times, a short bench line, the ncu launch list of that bench command and one
# `ncu --set full` capture of the conv kernels of a forward.  Outputs under gpurun_out/$1_*.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
timeout 300 python tools/layer_times.py 2 > gpurun_out/${tag}_layers.log 2>&1; echo "layers rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -c 24 -f -o gpurun_out/${tag}_conv \
  python tools/one_forward.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_layers.log | grep prof
