#!/bin/bash
# N-GPU round for the exchange: correctness of the peer form (bulk copies) and the default form, then the timing sweep.
N=${N:-2}
run() { timeout ${TMO:-90} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 "$@" 2>&1 | tail -${TAIL:-1}; }
echo "check p2p (bulk):"; LP_CHECK_MODE=p2p run tools/exchange_check.py
if [ -z "$SKIP_DEFAULT" ]; then echo "check default:"; run tools/exchange_check.py; fi
for g in ${CTAS:-0 148}; do echo "bench ctas $g:"; LP_EXCHANGE_CTAS=$g run tools/allreduce_bench.py; done
