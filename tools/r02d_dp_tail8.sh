#!/bin/bash
# 8-GPU version of tools/r02d_dp_tail.sh, one pass
n=8
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 20 --warmup 5 --no-eager --no-extra --no-cpu-baseline > gpurun_out/dp_tail.json 2> gpurun_out/dp_tail.err || tail -3 gpurun_out/dp_tail.err
  python -c "
import json; d=json.load(open('gpurun_out/dp_tail.json')); print('$*', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', (d.get('exchange') or {}).get('mode'))"
}
run OFB_DP_OVERLAP=1 OFB_DP_BLOCKS_PER_BUCKET=12 OFB_DP_TAIL_BLOCKS=0
run OFB_DP_OVERLAP=0
run OFB_DP_OVERLAP=1 OFB_DP_BLOCKS_PER_BUCKET=6 OFB_DP_TAIL_BLOCKS=0
