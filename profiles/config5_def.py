"""BASELINE config 5 (sparse-gamma deep exponential family, TraceMeanField_ELBO, AdagradRMSProp)
at the named size: x [320, 4096] synthetic Poisson counts (the Olivetti CSV of
examples/sparse_gamma_def.py:218-221 is downloaded at run time in the reference; no network here),
widths 100/40/15, fp32, P particles vectorised (BASELINE: 256 over 8 GPUs = 32 per GPU; both are
timed on one GPU).  Prints one JSON line per (P, mode).

usage: python profiles/config5_def.py [steps] [P ...]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import models  # noqa: E402
import pyro_b200 as pyro  # noqa: E402
from pyro_b200 import _native  # noqa: E402
from pyro_b200.infer import SVI, TraceMeanField_ELBO  # noqa: E402
from pyro_b200.optim import AdagradRMSProp  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
Ps = [int(a) for a in sys.argv[2:]] or [32, 256]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
N, PX = 320, 4096
rate = torch.distributions.Gamma(0.5, 0.5).sample((N, PX)) * 2.0
x = torch.poisson(rate).to(dev)

for P in Ps:
    for mode in ("eager", "graph"):
        pyro.clear_param_store()
        m = models.SparseGammaDEF(PX, (100, 40, 15), device=dev, dtype=torch.float32)
        elbo = TraceMeanField_ELBO(num_particles=P, vectorize_particles=True, max_plate_nesting=1)
        elbo.capture_graph = mode == "graph"
        svi = SVI(m.model, m.guide, AdagradRMSProp({"eta": 4.5, "t": 0.1}), elbo)
        try:
            for _ in range(4):
                loss = svi.step(x)
            torch.cuda.synchronize(dev)
            n0 = _native.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = svi.step(x)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            terms = P * N * PX
            print(json.dumps({"config": "sparse_gamma_def N=320x4096 widths 100/40/15", "particles": P,
                              "mode": mode, "ms_per_step": round(ms, 3), "steps_per_sec": round(1e3 / ms, 2),
                              "poisson_terms_per_sec": round(terms / (ms * 1e-3), 1),
                              "loss": round(float(loss), 1),
                              "own_launches_per_step": (_native.launch_count() - n0) / steps,
                              "mem_GB": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2)}))
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"particles": P, "mode": mode, "error": repr(e)[:300]}))
