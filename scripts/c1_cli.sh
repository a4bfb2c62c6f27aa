#!/bin/bash
# the C1 drop-in wall clock (cold GPU, then beside a process that holds a context) and the launch list of the final code
O=gpurun_out
python -c "
import bench, json, torch
def run(tag):
    d=bench.c1_cli_wall_clock()
    print(tag, json.dumps(d, indent=1))
run('COLD')
torch.zeros(1, device='cuda'); torch.cuda.synchronize()
run('WARM (this process holds a context)')
" > $O/c1.json 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-ingest --no-init --no-cli"
$B > $O/l_plain.json 2> $O/l_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_final.csv $B > $O/l_ncu.log 2>&1
echo rc=$?
