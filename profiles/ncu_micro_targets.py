"""The micro-benchmark rows that sit below 60 % of the HBM peak, one launch each (for `ncu --set full`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyro_b200.distributions as dist  # noqa: E402

dev = "cuda"
M = 1 << 26
torch.manual_seed(0)
pos = torch.rand(M, device=dev) * 2 + 0.5
pos2 = torch.rand(M, device=dev) * 2 + 0.5
unit = torch.rand(M, device=dev).clamp(0.01, 0.99)
cnt = torch.poisson(pos * 2)
x = torch.randn(M, device=dev).abs() + 0.01
for _ in range(2):
    dist.Beta(pos.requires_grad_(True), pos2)._fused_sum(unit, None, 1.0, 1.0, 1.0, True)
    dist.Gamma(pos.requires_grad_(True), pos2)._fused_sum(x, None, 1.0, 1.0, 1.0, True)
    dist.Poisson(pos.requires_grad_(True))._fused_sum(cnt, None, 1.0, 1.0, 1.0, True)
    pos.requires_grad_(False)
    K = 64
    rows = M // K
    conc = torch.rand(rows, K, device=dev) * 2 + 0.3
    v = torch.distributions.Dirichlet(torch.ones(K, device=dev)).sample((1024,)).repeat(rows // 1024, 1).clamp(min=1e-6)
    dist.Dirichlet(conc).log_prob(v)
    logits = torch.randn(rows, K, device=dev)
    idx = torch.randint(0, K, (rows,), device=dev)
    dist.Categorical(logits=logits).log_prob(idx)
torch.cuda.synchronize()
