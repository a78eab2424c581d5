#!/bin/bash
# ncu evidence for the round (one GPU; every profiled command first runs once WITHOUT ncu and has to exit 0):
#   launch list of the default bench command, `--set full` captures of msv1_decode_kernel (C2) and of the two ScreenPressor
#   kernels (C3 = I frames, C4 at 64 streams = P frames).  Reports stay in gpurun_out/, summaries are extracted on the build box.
T=${1:-r02}; O=gpurun_out; mkdir -p $O
C2="python bench.py --steps 2 --warmup 3 --no-codecs --no-cpu-baseline --e2e-steps 0"
C3="python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
C4="python bench.py --workload c4 --streams 64 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$C2 > $O/${T}_plain_c2.log 2>&1 || { echo "plain C2 run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_ncu_launches_c2.csv $C2 > $O/${T}_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msv1_decode -s 4 -c 1 -f -o $O/${T}_msv1 $C2 > $O/${T}_ncu_c2_full.log 2>&1
$C3 > $O/${T}_plain_c3.log 2>&1 || { echo "plain C3 run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/${T}_ncu_launches_c3.csv $C3 > $O/${T}_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sp2_i_kernel -s 4 -c 1 -f -o $O/${T}_sp2_i $C3 > $O/${T}_ncu_c3_full.log 2>&1
$C4 > $O/${T}_plain_c4.log 2>&1 || { echo "plain C4 run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:sp2_p_kernel -s 40 -c 1 -f -o $O/${T}_sp2_p $C4 > $O/${T}_ncu_c4_full.log 2>&1
ls -la $O/${T}_*.ncu-rep
