"""Launch the dominant kernels of BASELINE config 2 a few times (for ncu / timing).
usage: python profiles/prof_kernels.py [site|glm|normal|adam|all] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyro_b200.distributions as dist  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
P, N, D = 64, 1_000_000, 32
torch.manual_seed(0)
y = (torch.rand(N, device=dev) < 0.3).float()


def timed(name, fn, bytes_):
    """Device time of the launch sequence: captured into a CUDA graph so the host-side wrapper
    cost (~50 us of Python per call) is not inside the CUDA-event interval; L2 flushed between
    replays, outside the interval."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    print("%-30s %8.3f ms  %8.1f GB/s (algorithmic %d MB)" % (name, ms, bytes_ / ms / 1e6, bytes_ / 1e6))


if which in ("site", "all"):
    logits = torch.randn(P, N, device=dev).requires_grad_(True)
    timed("bernoulli fused sum+grad", lambda: dist.Bernoulli(logits=logits)._fused_sum(y, None, 1.0, -1.0 / P, 1.0, True),
          P * N * 8 + N * 4)
    lg2 = logits.detach()
    timed("bernoulli fused sum only", lambda: dist.Bernoulli(logits=lg2)._fused_sum(y, None, 1.0, 1.0, 1.0, True),
          P * N * 4 + N * 4)
if which in ("normal", "all"):
    M = 1 << 26
    x, loc = torch.randn(M, device=dev), torch.randn(M, device=dev)
    sc = torch.rand(M, device=dev) + 0.5
    timed("normal fused sum (3 operands)", lambda: dist.Normal(loc, sc)._fused_sum(x, None, 1.0, 1.0, 1.0, True), M * 12)
    locg = loc.clone().requires_grad_(True)
    timed("normal sum + dloc full", lambda: dist.Normal(locg, sc)._fused_sum(x, None, 1.0, 1.0, 1.0, True), M * 16)
    a = torch.rand(M, device=dev) * 3 + 0.2
    timed("gamma fused sum", lambda: dist.Gamma(a, sc)._fused_sum(x.abs() + 0.01, None, 1.0, 1.0, 1.0, True), M * 12)
if which in ("glm", "all"):
    X = torch.randn(N, D, device=dev)
    w = (0.1 * torch.randn(P, 1, D, device=dev)).requires_grad_(True)
    b = torch.zeros(P, 1, device=dev, requires_grad=True)
    timed("glm fused (X,y once)", lambda: dist.Bernoulli(logits=dist.linear_predictor(X, w, b))._fused_sum(y, None, 1.0, -1.0 / P, 1.0, True),
          N * D * 4 + N * 4)
torch.cuda.synchronize()
