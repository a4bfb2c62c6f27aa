#!/bin/bash
# end-of-round bench line on N GPUs of one box (peer all-reduce, multi_gpu_parity, strong-scaling configs)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu \
  > gpurun_out/z_n$N.json 2> gpurun_out/z_n$N.err
echo rc=$?
tail -c 600 gpurun_out/z_n$N.err
