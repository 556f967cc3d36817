#!/bin/bash
# One GPU call: a short bench line, the ncu launch list of that bench command, and `ncu --set full` captures of (a) the conv
# kernels, (b) the elementwise kernels of one forward, (c) the fused step kernels.  The .ncu-rep files are exported to CSV
# ON THE BOX (raw metrics per launch + cuda/sass source pages of the warm launches) and deleted: gpurun_out/ only travels
# back when it is under 64 MiB.  Outputs under gpurun_out/$1_*.   usage: tools/r02_profile.sh TAG [pytest]
tag=${1:-r02}
mkdir -p gpurun_out
if [ "$2" = "pytest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
fi
timeout 300 python tools/layer_times.py 2 > gpurun_out/${tag}_layers.log 2>&1; echo "layers rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
export_rep() {   # $1 = rep stem, $2.. = launch indices whose source page is exported
  stem=$1; shift
  ncu -i gpurun_out/${stem}.ncu-rep --page raw --csv > gpurun_out/${stem}_raw.csv 2>/dev/null
  for k in "$@"; do
    ncu -i gpurun_out/${stem}.ncu-rep --page source --print-source cuda,sass --csv --launch-skip $k --launch-count 1 > gpurun_out/${stem}_src_k$k.csv 2>/dev/null
  done
  rm -f gpurun_out/${stem}.ncu-rep
}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -c 20 -f -o gpurun_out/${tag}_conv \
  python tools/one_forward.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu conv rc=$?"
export_rep ${tag}_conv 10 11 13 16 17 18 19
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"upcat|maxpool|init_conv|gn_silu|temb" -c 16 -f -o gpurun_out/${tag}_ew \
  python tools/one_forward.py > gpurun_out/${tag}_ncu_ew.log 2>&1; echo "ncu ew rc=$?"
export_rep ${tag}_ew 9 10 15
timeout 300 python tools/bench_steps.py > gpurun_out/${tag}_steps.log 2>&1; echo "steps rc=$?"; cut -c1-160 gpurun_out/${tag}_steps.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|step_sde" -c 24 -f -o gpurun_out/${tag}_steps \
  env STEPS_REPS=1 python tools/bench_steps.py > gpurun_out/${tag}_ncu_steps.log 2>&1; echo "ncu steps rc=$?"
export_rep ${tag}_steps 11 13 15
grep prof gpurun_out/${tag}_layers.log
du -sh gpurun_out
