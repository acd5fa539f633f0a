#!/bin/bash
# A/B of the gradient-exchange strategies at N GPUs (default 8): one bench.py run per configuration, JSON lines into gpurun_out/.
#   bash tools/ddp_ab.sh [N]
N=${1:-8}
run() {  # name, extra env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r2_ddp${N}_$name.json 2> gpurun_out/r2_ddp${N}_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_ddp${N}_$name.json').read().strip().splitlines()[-1])
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value'], 1), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('no result', e)
PY
)"
}
run overlap32 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT -- --bucket-mb 32
run single_fp32 X=1 -- --bucket-mb 100000
run single_bf16 B200VIT_DDP_BF16=1 -- --bucket-mb 100000
run overlap128 X=1 -- --bucket-mb 128
grep -i "nvls" gpurun_out/r2_ddp${N}_overlap32.err | head -5
