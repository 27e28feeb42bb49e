#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun; outputs in gpurun_out/, summaries copied to profiles/).
#   bash tools/run_profiles.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# every program runs once without ncu first (must exit 0), then under ncu
python bench.py --steps 100 --warmup 3 --skip-cpu --skip-large --skip-sweep > $O/${R}_bench_plain.json 2> $O/${R}_bench_plain.err || exit 1
$NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/${R}_launches_bench_geballe_with_diamond.csv \
  python bench.py --steps 100 --warmup 3 --skip-cpu --skip-large --skip-sweep > $O/${R}_ncu_bench.log 2>&1
python tools/dev_ncu_target.py geballe_with_diamond 1.0 8 0 > $O/${R}_target_141k.log 2>&1 || exit 1
$NCU --set full --import-source on -k regex:k_pcg_pipe -s 5 -c 1 -o $O/${R}_k_pcg_pipe_141k -f \
  python tools/dev_ncu_target.py geballe_with_diamond 1.0 8 0 > $O/${R}_ncu_pipe.log 2>&1
python tools/dev_ncu_target.py konopkova 0.35 2 0 > $O/${R}_target_1m.log 2>&1 || exit 1
$NCU --set full --import-source on -k regex:k_pcg_stream -s 1 -c 1 -o $O/${R}_k_pcg_stream_1m -f \
  python tools/dev_ncu_target.py konopkova 0.35 2 0 > $O/${R}_ncu_stream_1m.log 2>&1
python tools/dev_ncu_target.py konopkova 0.18 2 0 rows > $O/${R}_target_4m.log 2>&1 || exit 1
$NCU --set full -k regex:k_pcg_stream -s 1 -c 1 -o $O/${R}_k_pcg_stream_4m -f \
  python tools/dev_ncu_target.py konopkova 0.18 2 0 rows > $O/${R}_ncu_stream_4m.log 2>&1
for f in $O/${R}_k_*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.csv 2>/dev/null; done
ls -la $O
