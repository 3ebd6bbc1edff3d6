# A/B of the z-chunk rule on N GPUs: the default against DGB_FV_ZCHUNK=128 (the rule of the sessions before), same box
N=$1; STEPS=${2:-100}
for z in auto 128 auto; do
  if [ $z = auto ]; then unset DGB_FV_ZCHUNK; else export DGB_FV_ZCHUNK=$z; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 10 --no-cpu --no-e2e > gpurun_out/ab${N}_$z.json 2> gpurun_out/ab${N}_$z.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/ab${N}_$z.json').read().strip().splitlines()[-1]); print('$z', d['ms_per_step'], d['roofline']['frac'], d['kernel'], d.get('parity_check',{}).get('ok'), d['halo']['exchange_ms'])"
done
