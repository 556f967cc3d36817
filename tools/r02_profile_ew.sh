#!/bin/bash
# Reduced profiling pass (one GPU call): the ncu launch list of the bench command and an `ncu --set full` capture of the
# elementwise kernels of one forward (init_conv_tc / maxpool / upcat / temb), raw metrics exported to CSV on the box.
# Outputs under gpurun_out/$1_*.   usage: tools/r02_profile_ew.sh TAG
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"upcat|maxpool|init_conv|gn_silu|temb" -c 16 -f -o gpurun_out/${tag}_ew \
  python tools/one_forward.py > gpurun_out/${tag}_ncu_ew.log 2>&1; echo "ncu ew rc=$?"
ncu -i gpurun_out/${tag}_ew.ncu-rep --page raw --csv > gpurun_out/${tag}_ew_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_ew.ncu-rep
du -sh gpurun_out
