#!/bin/bash
# A/B on ONE box: bench line (train step + Predictor, no CPU / cuDNN arms) of the tree in _ab_old/ (an older commit, built
# in-tree by the caller) and of the working tree, interleaved twice.  -> gpurun_out/ab.txt
# Prepare the old tree here (it travels with the gpurun snapshot, git-ignored):
#   mkdir _ab_old && git archive <commit> elektronn3_b200 include bench.py oracle/torch_ref.py oracle/__init__.py | tar -x -C _ab_old
#   (cd _ab_old && python -m elektronn3_b200.build)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
R=$(pwd)
: > gpurun_out/ab.txt
for rep in 1 2; do
  for tree in _ab_old .; do
    (cd $R/$tree && timeout 300 python bench.py --no-cpu-baseline --no-ref-gpu --steps 40 --warmup 10 2> /dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read())
p=b.get('predictor') or {}
print('$tree', 'train ms/step %.4f  e2e %.4f  roofline %.3f (%.1f us)  predictor %.4f s e2e %.4f s' % (b['ms_per_step'], b['e2e']['ms_per_step'], b['roofline']['frac'], 1e3*b['roofline']['ms_per_launch'], p.get('seconds_per_volume', 0), (p.get('e2e') or {}).get('seconds_per_volume', 0)))") | tee -a gpurun_out/ab.txt
  done
done
