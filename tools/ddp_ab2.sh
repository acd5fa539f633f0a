#!/bin/bash
# A/B of NCCL algorithm / protocol choices for the gradient all-reduce at N GPUs (default 8); JSON lines into gpurun_out/.
N=${1:-8}
run() {
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r2_nccl${N}_$name.json 2> gpurun_out/r2_nccl${N}_$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d = [json.loads(l) for l in open('gpurun_out/r2_nccl${N}_$name.json') if l.startswith('{')][-1]
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value'], 1), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('no result', e)
PY
)"
}
if [ "$2" = "diag" ]; then
  if [ "$3" = "avgsum" ]; then
    run sum_a X=1 --
    run avg_a B200VIT_DDP_AVG=1 --
    run sum_b X=1 --
    run avg_b B200VIT_DDP_AVG=1 --
    exit 0
  fi
  if [ "$3" = "sum" ]; then
    run sum_default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=TUNING --
    run sum_nvls "NCCL_ALGO=allreduce:NVLS" --
    run sum_fp32_nvls "NCCL_ALGO=allreduce:NVLS" -- --grad-compress none
    exit 0
  fi
  run skip_comm B200VIT_DDP_DIAG_SKIP_COMM=1 --
  run nvls_allreduce "NCCL_ALGO=allreduce:NVLS" --
  run default2 X=1 --
  exit 0
fi
run default X=1 --
run simple NCCL_PROTO=Simple --
run nvls NCCL_ALGO=NVLS --
run nvls_simple NCCL_ALGO=NVLS NCCL_PROTO=Simple --
run overlap32_simple NCCL_PROTO=Simple -- --bucket-mb 32 --grad-compress none
run single_fp32_simple NCCL_PROTO=Simple -- --grad-compress none
