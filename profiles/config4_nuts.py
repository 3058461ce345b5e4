"""BASELINE config 4's model (hierarchical Normal, J = 1e6 groups) through lockstep NUTS on the
fused leaf kernel: chain-leapfrogs/s for a bounded sample (few transitions, capped tree depth).
usage: python profiles/config4_nuts.py [chains] [warmup] [samples] [max_tree_depth]
       ncu --metrics gpu__time_duration.sum --csv --log-file L.csv python profiles/config4_nuts.py 32 2 2 5"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyro_b200.infer import MCMC, NUTS  # noqa: E402
from pyro_b200.infer.mcmc import HierNormalPotential  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = int(sys.argv[2]) if len(sys.argv) > 2 else 6
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 6
fused = (sys.argv[5] != "generic") if len(sys.argv) > 5 else True
dev = torch.device("cuda", 0)
J = 1_000_000
g = torch.Generator().manual_seed(0)
sig = (5 + 15 * torch.rand(J, generator=g)).to(dev)
yy = (5 + 3 * torch.randn(J, generator=g)).to(dev) + sig * torch.randn(J, generator=g).to(dev)
k = NUTS(potential_fn=HierNormalPotential(yy, sig, 10.0, 25.0), native_small=False, max_tree_depth=depth,
         fused_leaf=fused)
import time  # noqa: E402
marks = []


def hook(kernel, z, stage, t):
    torch.cuda.synchronize(dev)
    marks.append((stage, t, time.perf_counter(), kernel.leapfrog_count()))


mc = MCMC(k, num_samples=S, warmup_steps=W, num_chains=C, seed=0, hook_fn=hook if os.environ.get("B2_TRACE") else None)
torch.cuda.synchronize(dev)
t_start = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
mc.run()
e1.record()
e1.synchronize()
n = k.leapfrog_count()
sec = e0.elapsed_time(e1) * 1e-3
if marks:
    prev_t, prev_n = t_start, 0
    for stage, t, tm, nl in marks:
        print("%s %d: %.1f ms, %d chain-leapfrogs (%.1f k/s)" % (stage, t, (tm - prev_t) * 1e3, nl - prev_n,
                                                                 (nl - prev_n) / max(tm - prev_t, 1e-9) / 1e3))
        prev_t, prev_n = tm, nl
print(json.dumps({"config": "hier_normal J=1e6", "chains": C, "transitions": W + S, "max_tree_depth": depth,
                  "fused_leaf": fused, "chain_leapfrogs": n, "seconds": round(sec, 3),
                  "chain_leapfrog_per_sec": round(n / sec, 1),
                  "algorithmic_GBps_16B": round(n * 16e6 / sec / 1e9, 1),
                  "frac_of_410k_roofline": round(n / sec / 410e3, 3)}))
