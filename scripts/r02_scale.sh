#!/bin/bash
# cfg3 strong scaling at N = 2, 4, 8 (and cfg2 weak at 8) on one box, launched as the driver does
O=gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > $O/r02_scale_n$N.json 2> $O/r02_scale_n$N.err
  echo "N=$N rc=$?"; tail -c 600 $O/r02_scale_n$N.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config cfg2 --no-cpu-baseline > $O/r02_scale_cfg2_n8.json 2> $O/r02_scale_cfg2_n8.err
python bench.py --no-cpu-baseline > $O/r02_scale_n1.json 2>/dev/null
