#!/bin/bash
# Profile ONE rank of a multi-process run (developer tool): rank 0 runs under ncu with a metric
# set that fits ONE pass (a persistent kernel that talks to its neighbours cannot be replayed),
# the other ranks run normally.
#   python -m torch.distributed.run --no-python --nproc-per-node K tools/ncu_rank0.sh <metrics> <out.csv> <skip> script.py args...
METRICS="$1"; OUT="$2"; SKIP="$3"; shift 3
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics "$METRICS" --clock-control none --replay-mode kernel -k regex:world_kernel -s "$SKIP" -c 1 \
       --csv --log-file "$OUT" python "$@"
else
  exec python "$@"
fi
