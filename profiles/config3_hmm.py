"""BASELINE config 3 (GaussianHMM, H=512, O=4, T up to 10 000) timing: log_prob forward and
forward+backward w.r.t. all parameters.  usage: python profiles/config3_hmm.py [T] [tf32]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyro_b200.distributions as dist
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
tf32 = len(sys.argv) > 2 and sys.argv[2] == "tf32"
H, O = 512, 4
dev = "cuda"
torch.manual_seed(0)
F = (0.5 * torch.randn(H, H, device=dev) / H ** 0.5).requires_grad_(True)  # spectral radius ~0.5: stable
Hm = torch.randn(H, O, device=dev).requires_grad_(True)
tsc = (torch.randn(H, device=dev) * 0.1).exp().requires_grad_(True)
osc = (torch.randn(O, device=dev) * 0.1).exp().requires_grad_(True)
isc = torch.ones(H, device=dev).requires_grad_(True)
data = torch.randn(T, O, device=dev)
def build():
    return dist.GaussianHMM(dist.Normal(torch.zeros(H, device=dev), isc).to_event(1), F,
                            dist.Normal(torch.zeros(H, device=dev), tsc).to_event(1), Hm,
                            dist.Normal(torch.zeros(O, device=dev), osc).to_event(1), duration=T, tf32=tf32)
for what in ("fwd", "fwd+bwd"):
    torch.cuda.synchronize(); t0 = time.time()
    if what == "fwd":
        with torch.no_grad():
            lp = build().log_prob(data)
    else:
        lp = build().log_prob(data)
        lp.backward()
    torch.cuda.synchronize(); dt = time.time() - t0
    flops = T * (2 * 2 * H ** 3) * (1 if what == "fwd" else 3)
    print("T=%d %s: %.2f s  log_prob %.3f  (%.1f TFLOP/s on the H^3 GEMMs, tf32=%s, mem %.1f GB)" % (
        T, what, dt, float(lp), flops / dt / 1e12, tf32, torch.cuda.max_memory_allocated() / 1e9))
