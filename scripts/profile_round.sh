# ncu evidence for profiles/ (GPU box, one GPU): launch list of a short default bench + one --set full capture of the
# forward, BPTT and HierarchicalRNN per-coordinate kernels.  Usage: bash scripts/profile_round.sh <tag>
tag=${1:-r01b}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --coords 303104 --unroll 20 --no-cpu-baseline --no-also > gpurun_out/${tag}_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'unroll_(fwd|bwd)_kernel' -c 2 -f -o gpurun_out/${tag}_tc_fwd_bwd \
  python bench.py --steps 1 --warmup 0 --coords 303104 --unroll 20 --no-cpu-baseline --no-also > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'coord_kernel' -s 2 -c 1 -f -o gpurun_out/${tag}_hrnn_coord \
  python scripts/hrnn_profile.py > gpurun_out/${tag}_ncu_hrnn.log 2>&1
for r in ${tag}_tc_fwd_bwd ${tag}_hrnn_coord; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
ls -la gpurun_out/${tag}_*
