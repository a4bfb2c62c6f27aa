#!/bin/bash
# Round-2 profiling pass of the decode kernels (half-precision k_emis_dec, k_fwd_cells32) at the C4 shape, and the launch list
# of the bench command; CSV exports only (see profile_r2.sh).
set -x
F="--no-cpu --no-configs --no-ingest --no-init --no-cli"
O=gpurun_out
python bench.py --steps 2 --warmup 3 $F > $O/p_plain.json 2> $O/p_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 $F > $O/p_ncu_launch.log 2>&1
python scripts/c4_mid.py > $O/p_c4_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_emis_dec|k_fwd_cells32" --launch-skip 8 -c 2 -f -o /tmp/r2_dec python scripts/c4_mid.py > $O/p_ncu_dec.log 2>&1
ncu -i /tmp/r2_dec.ncu-rep --page raw --csv > $O/r2_dec_raw.csv
ncu -i /tmp/r2_dec.ncu-rep --page source --csv --print-source sass > $O/r2_dec_src.csv
cuobjdump -sass speech_recognition_hmm_continuous_b200/libhmmcu.so | grep -o "UTCHMMA[A-Z0-9_.]*\|UTCQMMA[A-Z0-9_.]*\|UBLKCP[A-Z0-9_.]*\|UTCBAR[A-Z0-9_.]*\|LDTM[A-Z0-9_.]*\|STTM[A-Z0-9_.]*" | sort | uniq -c > $O/r2_sass_mnemonics.txt
du -sh $O
