#!/bin/bash
# e2e of config 2 on N GPUs with one and with two pipelines per GPU
N=${1:-4}
for ctx in 2 1 2 1; do
  echo "== FNFT_B200_CTX_PER_DEVICE=$ctx, $N GPUs"
  FNFT_B200_CTX_PER_DEVICE=$ctx python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-extras --parity-signals 4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('  value %.0f ms/step %.2f  e2e %.0f ms/step %.2f %s ceiling %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['ms_each_step_rank0'],d['e2e'].get('copy_ceiling_gbs_all_ranks')))
"
done
