#!/bin/bash
# First GPU calls of a round, in the order that yields the most per GPU-minute (run each line through gpurun from /root/repo):
#
#  1 GPU  (~90 s):  tools/round_start.sh single
#  8 GPUs (~90 s):  tools/round_start.sh parity8      # strict multi-rank parity (not re-run at the end of round 1)
#  N GPUs         :  for n in 1 2 4 8; do tools/scale_run.sh $n takatak_b1025_t50; done
#  1 GPU  (ncu)   :  tools/round_start.sh ncu          # launch list of one bench step, after the plain run exited 0
set -u
mkdir -p gpurun_out
case "${1:-single}" in
  single)
    timeout 150 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
    timeout 150 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
    timeout 60 python tools/file_e2e.py > gpurun_out/file_e2e.json 2> gpurun_out/file_e2e.err; echo "file_e2e rc=$?"
    ;;
  parity8)
    PAMREC_TEST_RANKS=${2:-8} timeout 300 python -m pytest tests/test_gpu_sharded.py -q -x -s > gpurun_out/pytest_gpu${2:-8}.log 2>&1
    echo "rc=$?"; tail -15 gpurun_out/pytest_gpu${2:-8}.log
    ;;
  ncu)
    timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || exit 1
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
    echo "ncu rc=$?"
    ;;
esac
