#!/bin/bash
# ncu captures of k_accum_h and the tile packer k_pack_x16: C3-shard size (raw page, SASS source page) and C2 size (raw page)
set -x
O=gpurun_out
python scripts/dbg_acch.py 12500 > $O/h_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_accum_h|k_pack_x16" --launch-skip 1 -c 3 -f -o /tmp/r2h python scripts/dbg_acch.py 12500 > $O/h_ncu.log 2>&1
ncu -i /tmp/r2h.ncu-rep --page raw --csv > $O/r2h_raw.csv
ncu -i /tmp/r2h.ncu-rep --page source --csv --print-source sass -k k_accum_h > $O/r2h_src.csv
ncu --set full --clock-control none -k regex:"k_accum_h|k_pack_x16" --launch-skip 1 -c 3 -f -o /tmp/r2h2 python scripts/dbg_acch.py 1000 > $O/h_ncu_c2.log 2>&1
ncu -i /tmp/r2h2.ncu-rep --page raw --csv > $O/r2h_c2_raw.csv
du -sh $O
