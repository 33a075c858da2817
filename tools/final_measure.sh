#!/bin/sh
# Round-end measurements on the GPU box (one GPU): ncu captures first (never timed), then the bench lines.
#   gpurun --timeout 1800 -- 'sh tools/final_measure.sh r02'
R=${1:-r02}; O=gpurun_out; mkdir -p $O
ncu --set full --import-source on --clock-control none -k regex:realign_kernel -s 2 -c 1 -o $O/${R}_realign_final -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/${R}_ncu_realign.log 2>&1
python tools/make_traffic_json.py $O/${R}_realign_final.ncu-rep profiles/traffic.json $O/${R}_realign_kernel_ncu_full.csv && cp profiles/traffic.json $O/traffic.json
ncu --set full --import-source on --clock-control none -k regex:indel_support_ -c 3 -o $O/${R}_support_final -f \
    python bench.py --workload support --steps 1 --warmup 0 --no-cpu > $O/${R}_ncu_support.log 2>&1
ncu -i $O/${R}_support_final.ncu-rep --page raw --csv > $O/${R}_support_ncu_full.csv 2>/dev/null
python tools/ncu_summary.py $O/${R}_support_ncu_full.csv indel_support > $O/${R}_support_pack_ncu_summary.txt
python tools/ncu_summary.py $O/${R}_realign_kernel_ncu_full.csv realign_kernel > $O/${R}_realign_kernel_ncu_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err; tail -c 600 $O/${R}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err
python bench.py --numgaps 4 --reads 262144 --no-cpu --no-extra --steps 10 --warmup 3 > $O/${R}_bench_g4.json 2>/dev/null
head -c 700 $O/${R}_bench.json; echo; head -c 400 $O/${R}_bench_reference_arm.json; echo; head -c 300 $O/${R}_bench_g4.json
