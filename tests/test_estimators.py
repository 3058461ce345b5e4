"""SURVEY.md 8(f) row 4 -- other estimators on the same kernels: RenyiELBO against the reference's own
numbers (tests/golden/renyi.npz, recorded from unmodified Pyro by tests/golden/make_golden_r2.py) and
Predictive against a conjugate closed form.  CPU tier: host logic on the oracle-backed stand-ins; GPU tier:
the real kernels."""
import pytest
import torch
from torch.distributions import constraints

import pyro_b200 as pyro
import pyro_b200.distributions as dist
from conftest import EMULATE, device, load_npz
from pyro_b200 import poutine
from pyro_b200.infer import Predictive, RenyiELBO


def _renyi(dev):
    g = load_npz("renyi.npz")
    torch.set_default_dtype(torch.float64)
    X, y, eps = (torch.as_tensor(g[k]).to(dev) for k in ("X", "y", "eps"))
    N, D = X.shape
    P = int(g["P"])

    def model(X, y):
        w = pyro.sample("w", dist.Normal(torch.zeros(D, device=dev), torch.ones(D, device=dev)).to_event(1))
        with pyro.plate("data", N):
            mean = (X * w).sum(-1) if w.dim() == 1 else (w * X).sum(-1)
            pyro.sample("obs", dist.Normal(mean, torch.tensor(0.5, device=dev)), obs=y)

    class Inject(poutine.Messenger):
        def _pyro_sample(self, msg):
            if msg["name"] == "w" and not msg["is_observed"]:
                base = msg["fn"].base_dist
                msg["value"] = base.loc + eps * base.scale
                msg["done"] = True

    def guide(X, y):
        m = pyro.param("m", lambda: torch.tensor([0.1, -0.2, 0.3], device=dev))
        s = pyro.param("s", lambda: torch.tensor([0.5, 0.7, 0.9], device=dev), constraint=constraints.positive)
        with Inject():
            pyro.sample("w", dist.Normal(m, s).to_event(1))

    for alpha, tag in ((0.5, "0.5"), (2.0, "2")):
        pyro.clear_param_store()
        elbo = RenyiELBO(alpha=alpha, num_particles=P, vectorize_particles=True, max_plate_nesting=1)
        with poutine.trace(param_only=True) as cap:
            loss = elbo.loss_and_grads(model, guide, X, y)
        assert abs(loss - float(g["loss_" + tag])) <= 1e-9 * abs(float(g["loss_" + tag]))
        for name in ("m", "s"):
            got = cap.trace.nodes[name]["value"]._pyro_unconstrained_param.grad.cpu()
            assert torch.allclose(got, torch.as_tensor(g["grad_%s_%s" % (name, tag)]), atol=1e-9, rtol=1e-9), (alpha, name)
        assert abs(elbo.loss(model, guide, X, y) - float(g["value_" + tag])) <= 1e-9 * abs(float(g["value_" + tag]))


def _predictive(dev):
    """Normal-Normal: z ~ N(0, 1), x ~ N(z, 0.5).  Given posterior draws of z the predictive draws of x have
    mean E[z] and variance Var[z] + 0.25; both execution modes (sequential / one vectorised run) agree."""
    torch.set_default_dtype(torch.float32)
    torch.manual_seed(0)

    def model():
        z = pyro.sample("z", dist.Normal(torch.zeros((), device=dev), torch.ones((), device=dev)))
        with pyro.plate("d", 5):
            return pyro.sample("x", dist.Normal(z, torch.tensor(0.5, device=dev)))   # z: [] or [S, 1]

    S = 4000
    zs = 0.7 + 0.3 * torch.randn(S, device=dev)
    out = Predictive(model, posterior_samples={"z": zs}, parallel=True)()
    x = out["x"]
    assert tuple(x.shape) == (S, 5)
    assert abs(float(x.mean()) - 0.7) < 0.03
    assert abs(float(x.var()) - (0.09 + 0.25)) < 0.03
    seq = Predictive(model, posterior_samples={"z": zs[:50]})()
    assert tuple(seq["x"].shape) == (50, 5)

    def guide():
        pyro.sample("z", dist.Normal(torch.tensor(0.7, device=dev), torch.tensor(0.3, device=dev)))
    out2 = Predictive(model, guide=guide, num_samples=3000, parallel=True, return_sites=("x", "z"))()
    assert abs(float(out2["z"].mean()) - 0.7) < 0.03 and abs(float(out2["x"].var()) - 0.34) < 0.04


@pytest.fixture
def emu():
    import cpu_emulation
    with cpu_emulation.enabled():
        yield


def test_renyi_elbo_matches_reference_cpu(emu):
    _renyi("cpu")


def test_predictive_conjugate_cpu(emu):
    _predictive("cpu")


@pytest.mark.gpu
def test_renyi_elbo_matches_reference_gpu():
    if EMULATE:
        pytest.skip("covered by the cpu test")
    _renyi(device())


@pytest.mark.gpu
def test_predictive_conjugate_gpu():
    if EMULATE:
        pytest.skip("covered by the cpu test")
    _predictive(device())


def _tracegraph(dev):
    """TraceGraph_ELBO against unmodified Pyro (tests/golden/tracegraph.npz): nested plates, a Bernoulli and a
    Categorical guide site without rsample (decaying-average baseline / learnable baseline_value), a Normal site
    with rsample between them; guide values injected; two consecutive calls so the running baseline is used."""
    from pyro_b200.infer import TraceGraph_ELBO
    g = load_npz("tracegraph.npz")
    torch.set_default_dtype(torch.float64)
    t = lambda k: torch.as_tensor(g[k]).to(dev)  # noqa: E731
    data, a_val, b_val, eps, probs_b, qb0 = t("data"), t("a"), t("b"), t("eps"), t("probs_b"), t("qb0")
    c = lambda v: torch.tensor(v, device=dev)  # noqa: E731

    def model(data):
        with pyro.plate("outer", 3, dim=-1):
            a = pyro.sample("a", dist.Bernoulli(c(0.35)))
            z = pyro.sample("z", dist.Normal(2 * a - 1, c(1.0)))
            with pyro.plate("inner", 4, dim=-2):
                b = pyro.sample("b", dist.Categorical(probs_b[a.long()]))
                pyro.sample("obs", dist.Normal(z + b.to(data.dtype), c(1.5)), obs=data)

    class Inject(poutine.Messenger):
        def _pyro_sample(self, msg):
            if msg["name"] == "a":
                msg["value"] = a_val
            elif msg["name"] == "b":
                msg["value"] = b_val
            elif msg["name"] == "z":
                msg["value"] = msg["fn"].loc + eps * msg["fn"].scale

    def guide(data):
        qa = pyro.param("qa", lambda: c([0.4, 0.6, 0.5]), constraint=constraints.unit_interval)
        qb = pyro.param("qb", lambda: qb0.clone(), constraint=constraints.simplex)
        mz = pyro.param("mz", lambda: c([0.1, -0.3, 0.2]))
        sz = pyro.param("sz", lambda: c([0.8, 1.2, 0.6]), constraint=constraints.positive)
        bv = pyro.param("bv", lambda: torch.full((4, 3), -2.0, device=dev))
        with Inject(), pyro.plate("outer", 3, dim=-1):
            a = pyro.sample("a", dist.Bernoulli(qa),
                            infer={"baseline": {"use_decaying_avg_baseline": True, "baseline_beta": 0.8}})
            pyro.sample("z", dist.Normal(mz + a, sz))
            with pyro.plate("inner", 4, dim=-2):
                pyro.sample("b", dist.Categorical(qb), infer={"baseline": {"baseline_value": bv}})

    pyro.clear_param_store()
    elbo = TraceGraph_ELBO(max_plate_nesting=2)
    store = pyro.get_param_store()
    for it in range(2):
        loss = elbo.loss_and_grads(model, guide, data)
        assert abs(loss - float(g["loss_%d" % it])) <= 1e-9 * abs(float(g["loss_%d" % it]))
        for n in ("qa", "qb", "mz", "sz", "bv"):
            u = store._params[n]
            assert torch.allclose(u.grad.cpu(), torch.as_tensor(g["grad_%s_%d" % (n, it)]), atol=1e-9, rtol=1e-8), (n, it)
            u.grad = None
        avg = store._params["__baseline_avg_downstream_cost_a"].detach().cpu()
        assert torch.allclose(avg, torch.as_tensor(g["avg_a_%d" % it]), atol=1e-10, rtol=1e-10)
    assert abs(elbo.loss(model, guide, data) - float(g["value"])) <= 1e-9 * abs(float(g["value"]))
    torch.set_default_dtype(torch.float32)


def test_tracegraph_elbo_matches_reference_cpu(emu):
    _tracegraph("cpu")


@pytest.mark.gpu
def test_tracegraph_elbo_matches_reference_gpu():
    if EMULATE:
        pytest.skip("covered by the cpu test")
    _tracegraph(device())
