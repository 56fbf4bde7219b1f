#!/bin/bash
# 2-GPU runs of the default bench with pieces of the table gather switched off, to find what a
# second rank costs (under gpurun --gpus 2).  Logs: gpurun_out/r2k_<tag>.log
#   usage: bash profiles/n2_variants.sh [tag ...]      (default: all)
run() { tag=$1; lag=$2; shift 2; env "$@" IPB_BENCH_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 2 --steps 24 --warmup 8 --lag $lag \
        > gpurun_out/r2k_$tag.log 2>&1; }
want() { [ -z "$SEL" ] || [[ " $SEL " == *" $1 "* ]]; }
SEL="$*"
want base     && run base 2 X=1
want nogather && run nogather 2 IPB_DEBUG_NO_GATHER=1 IPB_DEBUG_NO_GATHER_D2H=1
want nod2h    && run nod2h 2 IPB_DEBUG_NO_GATHER_D2H=1
want nograph  && run nograph 2 IPB_GRAPHS=0
want lag1     && run lag1 1 X=1
true
