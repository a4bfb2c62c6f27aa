#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): launch list of the bench command and full captures of the hot kernels.
# The .ncu-rep files are exported to CSV on the box (raw page; source page of the decode kernels) and removed: gpurun brings
# back at most 64 MiB.
set -x
F="--no-cpu --no-configs --no-ingest --no-init --no-cli"
O=gpurun_out
python bench.py --steps 2 --warmup 3 $F > $O/p_plain.json 2> $O/p_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 $F > $O/p_ncu_launch.log 2>&1
python scripts/prof_estep.py 4 x > $O/p_estep_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"k_emis_ws|k_fb_res|k_accum_ws|k_finalize_slots|k_mstep" --launch-skip 10 -c 8 -f -o /tmp/r2_train python scripts/prof_estep.py 4 > $O/p_ncu_train.log 2>&1
ncu -i /tmp/r2_train.ncu-rep --page raw --csv > $O/r2_train_raw.csv
python scripts/c4_mid.py > $O/p_c4_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_emis_dec|k_fwd_cells32" --launch-skip 8 -c 2 -f -o /tmp/r2_dec python scripts/c4_mid.py > $O/p_ncu_dec.log 2>&1
ncu -i /tmp/r2_dec.ncu-rep --page raw --csv > $O/r2_dec_raw.csv
ncu -i /tmp/r2_dec.ncu-rep --page source --csv --print-source sass > $O/r2_dec_src.csv
ncu --set full --clock-control none -k regex:"k_emis_ws|k_fb_res|k_accum_ws" --launch-skip 9 -c 3 -f -o /tmp/r2_c3shard python bench.py --workload c3 --steps 2 --warmup 3 $F > $O/p_ncu_c3.log 2>&1
ncu -i /tmp/r2_c3shard.ncu-rep --page raw --csv > $O/r2_c3shard_raw.csv
du -sh $O
