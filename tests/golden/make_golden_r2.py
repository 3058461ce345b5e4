"""Round-2 additions to the golden fixtures (same recipe as make_golden.py: the UNMODIFIED reference imported
from /root/reference + the opt_einsum stand-in; run in the build container):

    python tests/golden/make_golden_r2.py

  renyi.npz    RenyiELBO (alpha = 0.5 and 2.0, 4 vectorised particles) loss and parameter gradients on a
               Normal regression with injected guide noise (pyro/infer/renyi_elbo.py)
  tracegraph.npz  TraceGraph_ELBO (pyro/infer/tracegraph_elbo.py): two consecutive loss_and_grads calls on a model with
               nested plates, two non-reparameterisable sites (decaying-average baseline / baseline_value), one
               reparameterised site, injected guide values; losses, parameter gradients, baseline state
  hmm_filter.npz  GaussianHMM.filter (pyro/distributions/hmm.py:604-633): posterior mean / covariance of the final
               hidden state, time-invariant and time-varying parameters
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "opt_einsum_standin"))
sys.path.insert(0, "/root/reference")

import pyro  # noqa: E402
import pyro.distributions as dist  # noqa: E402
import pyro.poutine as poutine  # noqa: E402
from pyro.infer import RenyiELBO  # noqa: E402

assert pyro.__version__ == "1.9.1"
torch.set_default_dtype(torch.float64)


def renyi():
    torch.manual_seed(3)
    N, D, P = 12, 3, 4
    X = torch.randn(N, D)
    y = X @ torch.tensor([0.5, -1.0, 0.25]) + 0.3 * torch.randn(N)
    eps = torch.randn(P, 1, D)
    out = {"X": X.numpy(), "y": y.numpy(), "eps": eps.numpy(), "P": P}

    def model(X, y):
        w = pyro.sample("w", dist.Normal(torch.zeros(D), torch.ones(D)).to_event(1))
        with pyro.plate("data", N):
            mean = (X * w).sum(-1) if w.dim() == 1 else (w * X).sum(-1)   # w: [P, 1, D] -> [P, N]
            pyro.sample("obs", dist.Normal(mean, 0.5), obs=y)

    class Inject(poutine.messenger.Messenger):
        def _pyro_sample(self, msg):
            if msg["name"] == "w" and not msg["is_observed"]:
                base = msg["fn"].base_dist
                msg["value"] = base.loc + eps * base.scale

    def guide(X, y):
        m = pyro.param("m", torch.tensor([0.1, -0.2, 0.3]))
        s = pyro.param("s", torch.tensor([0.5, 0.7, 0.9]), constraint=dist.constraints.positive)
        with Inject():
            pyro.sample("w", dist.Normal(m, s).to_event(1))

    for alpha in (0.5, 2.0):
        pyro.clear_param_store()
        elbo = RenyiELBO(alpha=alpha, num_particles=P, vectorize_particles=True, max_plate_nesting=1)
        loss = elbo.loss_and_grads(model, guide, X, y)
        store = pyro.get_param_store()
        out["loss_%g" % alpha] = loss
        out["grad_m_%g" % alpha] = store["m"].unconstrained().grad.numpy().copy()
        out["grad_s_%g" % alpha] = store["s"].unconstrained().grad.numpy().copy()
        out["value_%g" % alpha] = elbo.loss(model, guide, X, y)
    np.savez(os.path.join(HERE, "renyi.npz"), **out)
    print({k: (v if np.ndim(v) == 0 else np.asarray(v).shape) for k, v in out.items()})


def hmm_filter():
    torch.manual_seed(11)
    out = {}
    for tag, tv in (("inv", False), ("var", True)):
        Hd, O, T = 3, 2, 7
        lead = (T,) if tv else ()
        F = 0.5 * torch.randn(lead + (Hd, Hd))
        Hm = torch.randn(lead + (Hd, O))
        i_loc, i_scale = torch.randn(Hd), torch.rand(Hd) + 0.5
        t_loc, t_scale = torch.randn(lead + (Hd,)), torch.rand(lead + (Hd,)) + 0.3
        o_loc, o_scale = torch.randn(lead + (O,)), torch.rand(lead + (O,)) + 0.4
        x = torch.randn(T, O)
        hmm = dist.GaussianHMM(dist.Normal(i_loc, i_scale).to_event(1), F, dist.Normal(t_loc, t_scale).to_event(1),
                               Hm, dist.Normal(o_loc, o_scale).to_event(1), duration=T)
        post = hmm.filter(x)
        for k, v in (("F", F), ("H", Hm), ("i_loc", i_loc), ("i_scale", i_scale), ("t_loc", t_loc),
                     ("t_scale", t_scale), ("o_loc", o_loc), ("o_scale", o_scale), ("x", x),
                     ("mean", post.loc), ("cov", post.covariance_matrix), ("logp", hmm.log_prob(x))):
            out["%s.%s" % (tag, k)] = v.detach().numpy()
    np.savez(os.path.join(HERE, "hmm_filter.npz"), **out)
    print({k: np.asarray(v).shape for k, v in out.items()})


def tracegraph():
    from pyro.infer import TraceGraph_ELBO
    torch.manual_seed(5)
    data = torch.randn(4, 3)
    a_val = torch.tensor([1.0, 0.0, 1.0])
    b_val = torch.tensor([[0, 2, 1], [1, 1, 0], [2, 0, 2], [0, 1, 1]])
    eps = torch.randn(3)
    probs_b = torch.tensor([[0.2, 0.5, 0.3], [0.6, 0.1, 0.3]])
    qb0 = torch.softmax(torch.randn(4, 3, 3), -1)
    out = {"data": data.numpy(), "a": a_val.numpy(), "b": b_val.numpy(), "eps": eps.numpy(),
           "probs_b": probs_b.numpy(), "qb0": qb0.numpy()}

    def model(data):
        with pyro.plate("outer", 3, dim=-1):
            a = pyro.sample("a", dist.Bernoulli(torch.tensor(0.35)))
            z = pyro.sample("z", dist.Normal(2 * a - 1, 1.0))
            with pyro.plate("inner", 4, dim=-2):
                b = pyro.sample("b", dist.Categorical(probs_b[a.long()]))
                pyro.sample("obs", dist.Normal(z + b.to(data.dtype), 1.5), obs=data)

    class Inject(poutine.messenger.Messenger):
        def _pyro_sample(self, msg):
            if msg["name"] == "a":
                msg["value"] = a_val
            elif msg["name"] == "b":
                msg["value"] = b_val
            elif msg["name"] == "z":
                msg["value"] = msg["fn"].loc + eps * msg["fn"].scale

    def guide(data):
        qa = pyro.param("qa", torch.tensor([0.4, 0.6, 0.5]), constraint=dist.constraints.unit_interval)
        qb = pyro.param("qb", qb0, constraint=dist.constraints.simplex)
        mz = pyro.param("mz", torch.tensor([0.1, -0.3, 0.2]))
        sz = pyro.param("sz", torch.tensor([0.8, 1.2, 0.6]), constraint=dist.constraints.positive)
        bv = pyro.param("bv", torch.full((4, 3), -2.0))
        with Inject(), pyro.plate("outer", 3, dim=-1):
            a = pyro.sample("a", dist.Bernoulli(qa),
                            infer={"baseline": {"use_decaying_avg_baseline": True, "baseline_beta": 0.8}})
            pyro.sample("z", dist.Normal(mz + a, sz))
            with pyro.plate("inner", 4, dim=-2):
                pyro.sample("b", dist.Categorical(qb), infer={"baseline": {"baseline_value": bv}})

    pyro.clear_param_store()
    elbo = TraceGraph_ELBO(max_plate_nesting=2)
    names = ("qa", "qb", "mz", "sz", "bv")
    for it in range(2):
        loss = elbo.loss_and_grads(model, guide, data)
        store = pyro.get_param_store()
        out["loss_%d" % it] = loss
        for n in names:
            u = store[n].unconstrained()
            out["grad_%s_%d" % (n, it)] = u.grad.numpy().copy()
            u.grad = None
        out["avg_a_%d" % it] = store["__baseline_avg_downstream_cost_a"].detach().numpy().copy()
    out["value"] = elbo.loss(model, guide, data)
    np.savez(os.path.join(HERE, "tracegraph.npz"), **out)
    print({k: (v if np.ndim(v) == 0 else np.asarray(v).shape) for k, v in out.items()})


if __name__ == "__main__":
    renyi()
    hmm_filter()
    tracegraph()
