#!/bin/bash
# ncu --set full capture of one Poisson-solve launch (B=5: 15 clusters, one wave).  CHB_POISSON_V1=1 for the first generation.
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from ctrlhair_b200 import blend, synth
B = 5
cases = [synth.make_blend_case(256, 256, 900 + i) for i in range(B)]
face = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
gen = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
mask = 1 - blend.blend_mask(torch.from_numpy(np.stack([c[3] for c in cases])), torch.from_numpy(np.stack([c[2] for c in cases])))
out = blend.poisson_blending(face, gen, mask)
torch.cuda.synchronize()
PY
TAG=${1:-poisson}
timeout 280 ncu --set full --clock-control none --import-source on --kernel-name regex:poisson -c 1 -f -o gpurun_out/$TAG python /tmp/one.py > gpurun_out/ncu_$TAG.log 2>&1; echo "rc=$?"
