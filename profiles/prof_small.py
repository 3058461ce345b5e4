"""Driver for ncu captures of the one-CTA site kernels at the BASELINE config-2 latent-site shapes
([64, 1, 32] weights, [64, 1] bias): a few eager SVI steps of the logistic model on a small data set.
usage: ncu --set full --import-source on -k regex:site_small -c 12 python profiles/prof_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import models  # noqa: E402
from pyro_b200.infer import SVI, Trace_ELBO  # noqa: E402
from pyro_b200.optim import ClippedAdam  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
X = torch.randn(20000, 32, device=dev)
y = (torch.rand(20000, device=dev) < 0.5).float()
svi = SVI(models.logistic_model_fused, models.logistic_guide, ClippedAdam({"lr": 0.01}),
          Trace_ELBO(num_particles=64, vectorize_particles=True, max_plate_nesting=1))
for _ in range(3):
    print(svi.step(X, y))
