#!/bin/bash
# One-GPU evidence of a round, written to gpurun_out/<tag>_*: full GPU suite, bench lines of every workload, the reference arm,
# the ncu launch list of one bench step and the --set full capture of the step's kernels.   usage: tools/round_evidence.sh r02
set -u
T=${1:-r02}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${T}_pytest_gpu_final.log
timeout 300 python bench.py > gpurun_out/${T}_bench_final.json 2> gpurun_out/${T}_bench_final.err; echo "bench rc=$?"
for w in long_b4095_t200 wechat_b500_t100 eval_1p99 sharded_10m_b1025_t50; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; echo "bench $w rc=$?"
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
timeout 200 python tools/bench_siblings.py > gpurun_out/${T}_bench_siblings.jsonl 2> gpurun_out/${T}_bench_siblings.err; echo "siblings rc=$?"
timeout 60 python tools/file_e2e.py > gpurun_out/${T}_file_e2e.json 2> gpurun_out/${T}_file_e2e.err; echo "file_e2e rc=$?"
export PAMREC_GRAPH=0   # the profiler sees the step kernel by kernel (graph replay launches the same kernels)
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_plain.json 2> gpurun_out/${T}_bench_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 100 python tools/profile_step.py > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:'k_attn|k_proj|k_ffn|k_head2|k_embed_fwd|k_dtgt|k_sp2' --launch-skip 22 -c 22 \
    -o gpurun_out/${T}_step_full -f python tools/profile_step.py > gpurun_out/${T}_ncu_step_full.log 2>&1; echo "ncu full rc=$?"
