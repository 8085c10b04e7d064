#!/bin/bash
# usage: tools/dp_driver_check.sh N [extra driver flags]  -- the quick-start driver on synthetic text files under torchrun with N ranks, output to gpurun_out/
N=${1:-2}; shift
mkdir -p gpurun_out /tmp/dpchk
python - <<'PY'
from pamrec_b200 import synth
synth.generate("/tmp/dpchk/data", "wechat", n_users=400, n_items=3000, n_cates=40, mean_len=60, seed=7, eval_per_user=2)
PY
cd compat/example/00_quick_start
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29655 sequential.py \
  --dataset wechat --data_path /tmp/dpchk/data --epochs 1 --batch_size 100 --eval_step 5 --show_step 1000 --save_path /tmp/dpchk/out "$@" \
  > ../../../gpurun_out/dp_driver_$N.log 2>&1
echo "driver rc=$?"
tail -n 25 ../../../gpurun_out/dp_driver_$N.log
