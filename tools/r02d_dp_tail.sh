#!/bin/bash
# N-GPU A/B of the exchange placement: default (one all-reduce after the graph replay, AdamW from the host) against the
# block gradients' all-reduce launched inside the graph right after block 0's backward (overlaps the embed-stage tail), rest + AdamW
# in the graph. usage: tools/r02d_dp_tail.sh <ngpus>
n=${1:-2}
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 20 --warmup 5 --no-eager --no-extra --no-cpu-baseline > gpurun_out/dp_tail.json 2> gpurun_out/dp_tail.err || tail -3 gpurun_out/dp_tail.err
  python -c "
import json; d=json.load(open('gpurun_out/dp_tail.json')); print('$*', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', (d.get('exchange') or {}).get('mode'))"
}
for rep in 1 2; do
run OFB_DP_OVERLAP=0
run OFB_DP_OVERLAP=1 OFB_DP_BLOCKS_PER_BUCKET=12 OFB_DP_TAIL_BLOCKS=0
run OFB_DP_GRAPH=1
done
