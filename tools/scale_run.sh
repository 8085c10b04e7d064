#!/bin/bash
# usage: tools/scale_run.sh N [workload]   -- one bench line on N GPUs of this box, appended to gpurun_out/scale_<workload>.jsonl
N=$1; W=${2:-takatak_b1025_t50}
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 40 --warmup 5 --workload $W --no-cpu-baseline >> gpurun_out/scale_$W.jsonl 2>> gpurun_out/scale_$W.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + N)) bench.py --gpus $N --steps 40 --warmup 5 --workload $W >> gpurun_out/scale_$W.jsonl 2>> gpurun_out/scale_$W.err
fi
echo "N=$N $W rc=$?"
