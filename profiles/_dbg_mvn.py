import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import pyro_b200.distributions as dist
g = dict(np.load("/root/repo/tests/golden/dist_random.npz"))
dev="cuda"
v = torch.as_tensor(g["mvn_bcast.value"]).to(dev); loc = torch.as_tensor(g["mvn_bcast.p0"]).to(dev); L = torch.as_tensor(g["mvn_bcast.p1"]).to(dev)
ref = torch.as_tensor(g["mvn_bcast.lp"])
a = dist.MultivariateNormal(loc, scale_tril=L).log_prob(v).cpu()
b = dist.MultivariateNormal(loc.expand(19,4).contiguous(), scale_tril=L.expand(19,4,4).contiguous()).log_prob(v).cpu()
c = dist.MultivariateNormal(loc, scale_tril=L.expand(19,4,4).contiguous()).log_prob(v).cpu()
d = dist.MultivariateNormal(loc.expand(19,4).contiguous(), scale_tril=L).log_prob(v).cpu()
t = torch.distributions.MultivariateNormal(loc, scale_tril=L).log_prob(v).cpu()
print("shared  ", (a-ref).abs().max().item()); print("expanded", (b-ref).abs().max().item()); print("L exp   ", (c-ref).abs().max().item()); print("loc exp ", (d-ref).abs().max().item()); print("torch   ", (t-ref).abs().max().item())
print(L.stride(), loc.stride(), v.stride(), L.data_ptr()%16)
