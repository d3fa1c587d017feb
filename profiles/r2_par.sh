#!/bin/bash
# multi-GPU parity of all paths (one process per GPU).  usage: r2_par.sh N
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for P in resident tiles direct; do
  for D in bbm mevp; do      # (EVP over two steps is chaotic in the oracle itself)
    NSX_PATH=$P timeout 600 $TR tests/run_multigpu_parity.py --nx 128 --dyn $D --steps 2 2>&1 | grep -E "parity|MISMATCH|Error|error|timed out" | sed "s/^/$P /" | head -5
  done
done
