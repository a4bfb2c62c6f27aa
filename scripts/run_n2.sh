for pf in 8 128 64; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-configs --no-ingest --no-init --no-cli --set peer_fused=$pf > gpurun_out/r2h_n2_$pf.json 2> gpurun_out/r2h_n2.err; echo rc=$?
python -c "
import json,sys; s=open('gpurun_out/r2h_n2_$pf.json').read(); d=json.loads(s[s.index('{'):]); print('fused=$pf', d['ms_per_step'], d['e2e']['ms_per_step'], d['allreduce_ms'], d['allreduce_alone_ms'], d['multi_gpu_parity']['ok'])"
done
