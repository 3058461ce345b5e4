"""Host logic on the CPU tier: the native seams are replaced by oracle-backed stand-ins
(tests/cpu_emulation.py), everything else is the product's own Python -- effect handlers, plates,
Trace_ELBO assembly, optimiser bookkeeping, MCMC driver.  Results are compared with goldens from
the unmodified reference."""
import numpy as np
import pytest
import torch
from torch.distributions import constraints

import cpu_emulation
import models
import pyro_b200 as pyro
import pyro_b200.distributions as dist
from conftest import load_npz
from pyro_b200 import poutine
from pyro_b200.infer import SVI, Trace_ELBO, TraceMeanField_ELBO
from pyro_b200.optim import AdagradRMSProp, ClippedAdam


@pytest.fixture
def emu():
    with cpu_emulation.enabled():
        yield


def test_plate_shapes_and_scale(emu):
    def model():
        with pyro.plate("outer", 10, subsample_size=5, dim=-2):
            with pyro.plate("inner", 3, dim=-1):
                x = pyro.sample("x", dist.Normal(torch.tensor(0.0), torch.tensor(1.0)))
        return x
    tr = poutine.trace(model).get_trace()
    site = tr.nodes["x"]
    assert site["value"].shape == (5, 3)
    assert site["fn"].batch_shape == (5, 3)
    assert site["scale"] == 2.0
    assert [f.name for f in site["cond_indep_stack"]] == ["outer", "inner"]
    tr2 = poutine.prune_subsample_sites(tr)
    assert "outer" not in tr2.nodes


def test_plate_auto_dims_and_collision(emu):
    def model():
        with pyro.plate("a", 4):
            with pyro.plate("b", 2):
                return pyro.sample("x", dist.Normal(torch.zeros(()), torch.ones(())))
    assert poutine.trace(model).get_trace().nodes["x"]["value"].shape == (2, 4)
    with pytest.raises(ValueError):
        with pyro.plate("p", 3, dim=-1), pyro.plate("q", 3, dim=-1):
            pass


def test_replay_condition_block(emu):
    def model():
        z = pyro.sample("z", dist.Normal(torch.tensor(0.0), torch.tensor(1.0)))
        return pyro.sample("x", dist.Normal(z, torch.tensor(1.0)), obs=torch.tensor(0.3))
    g = poutine.trace(model).get_trace()
    r = poutine.trace(poutine.replay(model, trace=g)).get_trace()
    assert r.nodes["z"]["value"] is g.nodes["z"]["value"]
    c = poutine.trace(poutine.condition(model, data={"z": torch.tensor(1.5)})).get_trace()
    assert c.nodes["z"]["is_observed"] and float(c.nodes["z"]["value"]) == 1.5
    with poutine.trace() as outer:
        poutine.block(model, hide=["z"])()
    assert "z" not in outer.trace.nodes and "x" in outer.trace.nodes
    lp = c.log_prob_sum()
    ref = dist_ref_normal(1.5, 0, 1) + dist_ref_normal(0.3, 1.5, 1)
    assert abs(float(lp) - ref) < 1e-6


def dist_ref_normal(x, m, s):
    return float(torch.distributions.Normal(m, s).log_prob(torch.tensor(x)))


@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-9), ("f32", torch.float32, 3e-4)])
@pytest.mark.parametrize("elbo_cls", [Trace_ELBO])
@pytest.mark.parametrize("fused_draw", [False, True])
def test_svi_logistic_matches_reference_trajectory(emu, monkeypatch, tag, dtype, tol, elbo_cls, fused_draw):
    """The full host stack (particle plate, replay, fused-site ELBO assembly, per-parameter
    ClippedAdam state) reproduces the reference's 5-step SVI trajectory -- with the guide draws
    scored site by site, and with draw + score claimed from the fused rsample node."""
    g = load_npz("svi_logistic.npz")
    torch.set_default_dtype(dtype)
    X, y = torch.as_tensor(g["X"]).to(dtype), torch.as_tensor(g["y"]).to(dtype)
    eps_w, eps_b = torch.as_tensor(g["eps_w"]).to(dtype), torch.as_tensor(g["eps_b"]).to(dtype)
    P = int(g["P"])
    box = {"i": 0}

    def guide(X, y):
        with models.InjectNoise({"w": eps_w[box["i"]], "b": eps_b[box["i"]]}, fused_draw=fused_draw):
            models.logistic_guide(X, y)

    svi = SVI(models.logistic_model, guide, ClippedAdam({"lr": 0.01}),
              elbo_cls(num_particles=P, vectorize_particles=True, max_plate_nesting=1))
    claims = []
    real_claim = dist.claim_rsample_score
    monkeypatch.setattr(dist, "claim_rsample_score",
                        lambda fn, value, coeff: claims.append(real_claim(fn, value, coeff)) or claims[-1])
    for i in range(eps_w.shape[0]):
        box["i"] = i
        loss = svi.step(X, y)
        # both guide sites (and only those) hand over a precomputed sum log q when the draw is fused
        assert sum(c is not None for c in claims) == (2 * (i + 1) if fused_draw else 0)
        assert abs(loss - g["losses_" + tag][i]) <= 10 * tol * abs(g["losses_" + tag][i]), i
        store = pyro.get_param_store()
        flat = torch.cat([store[k].detach().reshape(-1).double() for k in ("w_loc", "w_scale", "b_loc", "b_scale")])
        assert torch.allclose(flat, torch.as_tensor(g["params_" + tag][i]), atol=10 * tol, rtol=10 * tol), i


@pytest.mark.parametrize("cls,tag", [(Trace_ELBO, "trace"), (TraceMeanField_ELBO, "meanfield")])
def test_elbo_grads_gamma_poisson_mask_subsample(emu, cls, tag):
    g = load_npz("elbo_grad.npz")
    torch.set_default_dtype(torch.float64)
    data, counts = torch.as_tensor(g["data"]), torch.as_tensor(g["counts"])
    mask = torch.as_tensor(g["mask"])
    eps, ueps = torch.as_tensor(g["eps"]), torch.as_tensor(g["ueps"])
    N = data.shape[0]

    def model():
        z = pyro.sample("z", dist.Normal(torch.tensor(0.0), torch.tensor(2.0)))
        rate = pyro.sample("rate", dist.Gamma(torch.tensor(2.0), torch.tensor(0.5)))
        with pyro.plate("data", 2 * N, subsample_size=N, dim=-1):
            with poutine.mask(mask=mask):
                pyro.sample("x", dist.Normal(z, torch.tensor(1.3)), obs=data)
            pyro.sample("c", dist.Poisson(rate), obs=counts)

    class Inject(poutine.Messenger):
        def __init__(self, vals):
            self.vals = vals

        def _pyro_sample(self, msg):
            if msg["name"] in self.vals:
                msg["value"] = self.vals[msg["name"]]
                msg["done"] = True

    def guide():
        loc = pyro.param("loc", torch.tensor(0.3))
        scale = pyro.param("scale", torch.tensor(0.7), constraint=constraints.positive)
        conc = pyro.param("conc", torch.tensor(3.0), constraint=constraints.positive)
        grate = pyro.param("grate", torch.tensor(1.2), constraint=constraints.positive)
        with Inject({"z": loc + eps * scale, "rate": conc / grate * (0.5 + ueps)}):
            pyro.sample("z", dist.Normal(loc, scale))
            pyro.sample("rate", dist.Gamma(conc, grate))

    elbo = cls(num_particles=6, vectorize_particles=True, max_plate_nesting=1)
    with poutine.trace(param_only=True) as cap:
        loss = elbo.loss_and_grads(model, guide)
    assert abs(loss - float(g[tag + ".loss"])) < 1e-8 * abs(float(g[tag + ".loss"]))
    for name, site in cap.trace.nodes.items():
        ref = torch.as_tensor(g["%s.grad.%s" % (tag, name)])
        got = site["value"]._pyro_unconstrained_param.grad
        assert torch.allclose(got, ref, atol=1e-7, rtol=1e-7), name


def test_score_function_path_matches_reference_kat(emu):
    """Non-reparameterised guide site -> general path with log_r (tests/infer/test_gradient.py
    style): compare with an independent autograd computation of the surrogate."""
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(0)
    data = torch.tensor([1.0, 0.0, 1.0, 1.0])
    zs = torch.tensor([[1.0], [0.0], [1.0]])  # 3 particles

    def model():
        p = pyro.sample("z", dist.Bernoulli(probs=torch.tensor(0.4)))
        with pyro.plate("d", 4):
            pyro.sample("x", dist.Bernoulli(probs=0.2 + 0.6 * p), obs=data)  # p: [P, 1] -> [P, 4]

    class Inject(poutine.Messenger):
        def _pyro_sample(self, msg):
            if msg["name"] == "z":
                msg["value"] = zs
                msg["done"] = True

    def guide():
        q = pyro.param("q", torch.tensor(0.3), constraint=constraints.unit_interval)
        with Inject():
            pyro.sample("z", dist.Bernoulli(probs=q))

    elbo = Trace_ELBO(num_particles=3, vectorize_particles=True, max_plate_nesting=1)
    with poutine.trace(param_only=True) as cap:
        loss = elbo.loss_and_grads(model, guide)
    got = cap.trace.nodes["q"]["value"]._pyro_unconstrained_param.grad
    # independent computation
    u = torch.tensor(0.3).logit().clone().requires_grad_(True)
    q = torch.sigmoid(u)
    B = torch.distributions.Bernoulli
    lq = B(probs=q).log_prob(zs)                        # [3,1]
    lpz = B(probs=torch.tensor(0.4)).log_prob(zs)
    lpx = B(probs=0.2 + 0.6 * zs).log_prob(data)         # [3,4]
    log_r = (lpz - lq).detach() + lpx.sum(-1, keepdim=True).detach()
    surrogate = -((log_r * lq).sum()) / 3
    surrogate.backward()
    assert torch.allclose(got, u.grad, atol=1e-10)
    elbo_val = (lpz + lpx.sum(-1, keepdim=True) - lq).sum() / 3
    assert abs(loss + float(elbo_val)) < 1e-10


def test_optimizer_checkpoint_roundtrip(emu, tmp_path):
    """tests/optim/test_optim.py:372-437: save -> clear -> load -> identical trajectory."""
    torch.set_default_dtype(torch.float64)
    g = load_npz("svi_logistic.npz")
    X, y = torch.as_tensor(g["X"]), torch.as_tensor(g["y"])
    eps_w, eps_b = torch.as_tensor(g["eps_w"]), torch.as_tensor(g["eps_b"])
    box = {"i": 0}

    def guide(X, y):
        with models.InjectNoise({"w": eps_w[box["i"]], "b": eps_b[box["i"]]}):
            models.logistic_guide(X, y)

    def make():
        return SVI(models.logistic_model, guide, ClippedAdam({"lr": 0.01, "lrd": 0.9}),
                   Trace_ELBO(num_particles=8, vectorize_particles=True, max_plate_nesting=1))
    svi = make()
    for i in range(2):
        box["i"] = i
        svi.step(X, y)
    svi.optim.save(str(tmp_path / "opt.pt"))
    pyro.get_param_store().save(str(tmp_path / "params.pt"))
    state = svi.optim.get_state()
    assert set(state) == {"w_loc", "w_scale", "b_loc", "b_scale"}
    assert state["w_loc"]["state"][0]["step"] == 2
    assert abs(state["w_loc"]["param_groups"][0]["lr"] - 0.01 * 0.9 ** 2) < 1e-15
    ref = []
    for i in range(2, 5):
        box["i"] = i
        ref.append(svi.step(X, y))
    pyro.clear_param_store()
    pyro.get_param_store().load(str(tmp_path / "params.pt"))
    svi2 = make()
    svi2.optim.load(str(tmp_path / "opt.pt"))
    got = []
    for i in range(2, 5):
        box["i"] = i
        got.append(svi2.step(X, y))
    assert np.allclose(ref, got, rtol=1e-12)


def test_adagrad_rmsprop_matches_reference(emu):
    g = load_npz("optim.npz")
    torch.set_default_dtype(torch.float64)
    p = torch.as_tensor(g["p0_f64"]).clone().requires_grad_(True)
    opt = AdagradRMSProp({"eta": 4.5, "t": 0.1})
    pyro.get_param_store()._param_to_name[p] = "p"
    for i, gr in enumerate(torch.as_tensor(g["grads_f64"])):
        p.grad = gr.clone()
        opt([p])
        assert torch.allclose(p.detach(), torch.as_tensor(g["adagrad_rmsprop_f64"][i]), atol=1e-12)


def test_lockstep_nuts_recovers_posterior_logistic(emu):
    """The lockstep iterative tree driver (host logic + masks) on a 3-d logistic regression:
    posterior means agree with the oracle's recursive NUTS within MC error
    (tests/infer/mcmc/test_nuts.py::test_logistic_regression tolerance style)."""
    from oracle import mcmc as omcmc
    from pyro_b200.infer import MCMC, NUTS
    from pyro_b200.infer.mcmc import LogisticPotential
    torch.set_default_dtype(torch.float64)
    g = load_npz("mcmc.npz")
    X, y = torch.as_tensor(g["lr.X"]), torch.as_tensor(g["lr.y"])
    kernel = NUTS(potential_fn=LogisticPotential(X, y, 1.0), native_small=False)
    mc = MCMC(kernel, num_samples=150, warmup_steps=100, num_chains=6, seed=1)
    mc.run()
    s = mc.get_samples()["beta"]
    chain = omcmc.NUTSChain(omcmc.logistic_potential(X, y, 1.0), 3, seed=2)
    ref, _ = chain.run(torch.zeros(3, dtype=torch.float64), 150, 600)
    assert torch.allclose(s.mean(0), ref.mean(0), atol=0.12)
    assert torch.allclose(s.std(0), ref.std(0), atol=0.08)
    d = mc.diagnostics()
    assert float(d["beta"]["r_hat"].max()) < 1.1
    assert kernel.leapfrog_count() > 0


def test_hmc_fixed_length_recovers_posterior_logistic(emu):
    """Fixed-length HMC (pyro/infer/mcmc/hmc.py:371-438: momentum draw, masked per-chain step counts,
    Metropolis correction, step-size / mass adaptation) on the 3-d logistic regression: posterior
    moments agree with the oracle's recursive NUTS within MC error."""
    from oracle import mcmc as omcmc
    from pyro_b200.infer import HMC, MCMC
    from pyro_b200.infer.mcmc import LogisticPotential
    torch.set_default_dtype(torch.float64)
    g = load_npz("mcmc.npz")
    X, y = torch.as_tensor(g["lr.X"]), torch.as_tensor(g["lr.y"])
    kernel = HMC(potential_fn=LogisticPotential(X, y, 1.0), step_size=0.1, trajectory_length=1.0)
    mc = MCMC(kernel, num_samples=200, warmup_steps=100, num_chains=6, seed=4)
    mc.run()
    s = mc.get_samples()["beta"]
    chain = omcmc.NUTSChain(omcmc.logistic_potential(X, y, 1.0), 3, seed=2)
    ref, _ = chain.run(torch.zeros(3, dtype=torch.float64), 150, 600)
    assert torch.allclose(s.mean(0), ref.mean(0), atol=0.12)
    assert torch.allclose(s.std(0), ref.std(0), atol=0.08)
    acc = torch.tensor(mc.diagnostics()["acceptance rate"])
    assert float(acc.min()) > 0.5


def test_fused_leaf_lockstep_driver_eight_schools(emu):
    """Host logic of the fused-leaf lockstep driver (``NUTS._sample_lockstep_hier``: per-depth
    merges, proposal flush, global-gradient bookkeeping) with the leaf kernel emulated: posterior of
    eight_schools against the reference's long run (goldens es.long.*)."""
    from pyro_b200.infer import MCMC, NUTS
    from pyro_b200.infer.mcmc import HierNormalPotential
    torch.set_default_dtype(torch.float64)
    g = load_npz("mcmc.npz")
    y, sigma = torch.as_tensor(g["es.y"]), torch.as_tensor(g["es.sigma"])
    kernel = NUTS(potential_fn=HierNormalPotential(y, sigma, 10.0, 25.0), native_small=False)
    mc = MCMC(kernel, num_samples=150, warmup_steps=100, num_chains=6, seed=3)
    mc.run()
    assert kernel._use_fused_hier
    s = mc.get_samples()
    assert abs(float(s["mu"].mean()) - float(g["es.long.mu.mean"][0])) < 1.0
    assert abs(float(s["tau"].mean()) - float(g["es.long.tau.mean"][0])) < 1.5
    assert float((s["eta"].mean(0) - torch.as_tensor(g["es.long.eta.mean"])).abs().max()) < 0.25
    assert kernel.leapfrog_count() > 6 * 250


def test_save_params_and_streaming_statistics(emu):
    """BASELINE config 4 cannot store its samples: with ``save_params`` only those sites are kept and
    every site's per-chain running mean / variance is maintained on the device.  Same seed, two runs:
    the kept sites equal the full run's, and the streaming statistics (per chain, and pooled over
    chains) equal the statistics of the full run's stored samples."""
    from pyro_b200.infer import MCMC, NUTS
    from pyro_b200.infer.mcmc import HierNormalPotential
    torch.set_default_dtype(torch.float64)
    g = load_npz("mcmc.npz")
    y, sigma = torch.as_tensor(g["es.y"]), torch.as_tensor(g["es.sigma"])

    def run(**kw):
        k = NUTS(potential_fn=HierNormalPotential(y, sigma, 10.0, 25.0), native_small=False)
        mc = MCMC(k, num_samples=30, warmup_steps=25, num_chains=3, seed=5, **kw)
        mc.run()
        return mc

    full = run()
    lean = run(save_params=["mu", "tau"])
    sf = full.get_samples(group_by_chain=True)
    sl = lean.get_samples(group_by_chain=True)
    assert set(sl) == {"mu", "tau"} and lean._samples.shape[-1] == 2
    for name in ("mu", "tau"):
        assert torch.equal(sl[name], sf[name])
    per_chain = lean.streaming_stats(pooled=False)
    pooled = lean.streaming_stats(pooled=True)
    for name in ("mu", "tau", "eta"):
        x = sf[name]                                   # [C, T, *shape]
        assert torch.allclose(per_chain[name]["mean"], x.mean(1), atol=1e-10)
        assert torch.allclose(per_chain[name]["variance"], x.var(1, unbiased=True), atol=1e-10)
        flat = x.reshape((-1,) + x.shape[2:])
        assert torch.allclose(pooled[name]["mean"], flat.mean(0), atol=1e-10)
        assert torch.allclose(pooled[name]["variance"], flat.var(0, unbiased=True), atol=1e-10)
        assert pooled[name]["n"] == flat.shape[0]


def test_trace_potential_matches_reference(emu):
    """Generic model potential (model run under a chain plate, fused site scoring) == reference
    potential + gradient at the golden points."""
    from pyro_b200.infer.mcmc import TracePotential
    torch.set_default_dtype(torch.float64)
    g = load_npz("mcmc.npz")
    y, sigma = torch.as_tensor(g["es.y"]), torch.as_tensor(g["es.sigma"])
    Z = torch.as_tensor(g["es.Z"])
    pot = TracePotential(models.eight_schools, (sigma, y), {}, num_chains=Z.shape[0])
    # golden layout [mu, log tau, eta]; TracePotential orders sites as the model declares them
    z = torch.cat([Z[:, pot_slice(pot, "mu", Z)], ], dim=1) if False else None
    order = list(pot.sites)
    cols = {"mu": Z[:, 0:1], "tau": Z[:, 1:2], "eta": Z[:, 2:]}
    zz = torch.cat([cols[n] for n in order], dim=1)
    U, G = pot.value_and_grad(zz)
    assert torch.allclose(U, torch.as_tensor(g["es.U"]), atol=1e-9, rtol=1e-9)
    ref_cols = {"mu": torch.as_tensor(g["es.G"])[:, 0:1], "tau": torch.as_tensor(g["es.G"])[:, 1:2],
                "eta": torch.as_tensor(g["es.G"])[:, 2:]}
    assert torch.allclose(G, torch.cat([ref_cols[n] for n in order], dim=1), atol=1e-9, rtol=1e-9)


def pot_slice(pot, name, Z):
    return pot.sites[name][0]


def test_stats_match_reference():
    from pyro_b200.infer.mcmc.stats import effective_sample_size, split_gelman_rubin
    g = load_npz("mcmc.npz")
    x = torch.as_tensor(g["stats.x"])
    assert torch.allclose(split_gelman_rubin(x), torch.as_tensor(g["stats.rhat"]), atol=1e-10)
    assert torch.allclose(effective_sample_size(x), torch.as_tensor(g["stats.neff"]), rtol=1e-8)


def test_vectorised_adaptation_matches_reference_pieces():
    from pyro_b200.infer.mcmc.adaptation import DualAveraging, WelfordDiag, build_adaptation_schedule
    import math
    g = load_npz("mcmc.npz")
    for w in (5, 19, 100, 150, 200, 500, 1000):
        assert [[a.start, a.end] for a in build_adaptation_schedule(w)] == g["sched.%d" % w].tolist()
    da = DualAveraging(2, torch.device("cpu"), prox_center=math.log(10 * 0.3))
    for gg, ref in zip(g["da.g"], g["da.x"]):
        da.step(torch.full((2,), float(gg), dtype=torch.float64))
        xt, xavg = da.get_state()
        assert abs(float(xt[1]) - ref[0]) < 1e-12 and abs(float(xavg[0]) - ref[1]) < 1e-12
    wf = WelfordDiag()
    for s in torch.as_tensor(g["wf.samples"]):
        wf.update(s.expand(3, -1))
    assert torch.allclose(wf.get_covariance(True)[2], torch.as_tensor(g["wf.cov_reg"]), atol=1e-12)


def _def_meanfield(device):
    g = load_npz("def_meanfield.npz")
    torch.set_default_dtype(torch.float64)
    x = torch.as_tensor(g["x"]).to(device)
    P, widths = int(g["P"]), tuple(int(w) for w in g["widths"])
    inj = {s_: (lambda a, r, k=k: (a / r) * (0.6 + 0.1 * k)) for k, s_ in
           enumerate(["w_top", "w_mid", "w_bottom", "z_top", "z_mid", "z_bottom"])}
    m = models.SparseGammaDEF(x.shape[1], widths, device=device, dtype=torch.float64, inject=inj, particles=P)
    svi = SVI(m.model, m.guide, AdagradRMSProp({"eta": 0.5, "t": 0.1}),
              TraceMeanField_ELBO(num_particles=P, vectorize_particles=True, max_plate_nesting=1))
    losses = [svi.step(x) for _ in range(6)]
    assert np.allclose(losses, g["losses"], rtol=1e-9), (losses, g["losses"])
    for k, v in pyro.get_param_store().named_parameters():
        assert torch.allclose(v.detach().cpu(), torch.as_tensor(g["param." + k]), atol=1e-9, rtol=1e-9), k


def test_sparse_gamma_def_meanfield_matches_reference(emu):
    """BASELINE config 5 structure (three Gamma layers + Poisson likelihood, TraceMeanField_ELBO with
    Gamma||Gamma KL terms, AdagradRMSProp): 6-step loss trajectory and final parameters of the
    unmodified reference."""
    _def_meanfield("cpu")


def test_fused_draw_graph_is_released(emu):
    """The fused draw's autograd node must die with the draw (a tag <-> node reference cycle would
    run through the C++ graph, which Python's GC cannot break: every step's graph, and the
    AccumulateGrad nodes of the parameters with it, would stay alive -- which also breaks CUDA-graph
    capture on a side stream)."""
    import gc
    import weakref
    loc = torch.zeros(3, requires_grad=True)
    scale = torch.ones(3, requires_grad=True)
    z = dist.Normal(loc, scale).rsample((5,))
    tag = z._b2_rsample
    assert tag.lq.requires_grad and tag.lq.shape == ()
    alive = weakref.ref(tag.coeff)
    del z, tag
    gc.collect()
    assert alive() is None


def test_optimizer_repoints_gradient_table_after_capture(emu, monkeypatch):
    """The captured SVI step lets autograd assign fresh gradient tensors; during capture the optimiser
    must keep its device table (its ADDRESS is baked into the captured launch) and only re-point the
    gradient pointers afterwards (``flush_pending``).  A changed parameter set during capture is refused."""
    from pyro_b200 import _native as N
    torch.set_default_dtype(torch.float64)
    p, q = torch.randn(5, requires_grad=True), torch.randn(3, requires_grad=True)
    store = pyro.get_param_store()
    store._param_to_name[p], store._param_to_name[q] = "p", "q"
    p.grad, q.grad = torch.ones(5), torch.ones(3)
    opt = ClippedAdam({"lr": 0.01})
    opt([p, q])
    table = opt._tables[torch.float64]
    g_addr = table["g"].data_ptr()
    monkeypatch.setattr(N, "capturing", lambda: True)
    p.grad, q.grad = torch.full((5,), 2.0), torch.full((3,), 2.0)      # what a grad-less backward leaves behind
    opt([p, q])
    assert opt._tables[torch.float64] is table and "pending" in table
    with pytest.raises(RuntimeError):
        opt([p])                                                       # different parameter set: refused
    monkeypatch.setattr(N, "capturing", lambda: False)
    opt.flush_pending()
    assert "pending" not in table and table["g"].data_ptr() == g_addr
    assert table["g"].tolist() == [p.grad.data_ptr(), q.grad.data_ptr()]
    assert table["key"] == opt._table_key([p, q])


def _full_mass_case(dev):
    """Dense mass matrix (pyro/infer/mcmc/adaptation.py:238-392, ``full_mass=True``) through the whitened
    coordinates of ``potential.WhitenedPotential``: on a posterior with correlation -0.99 between two
    coefficients the dense metric must (i) sample the same posterior as the oracle's recursive NUTS,
    (ii) learn a factor whose A A^T carries that correlation, (iii) settle at a much larger step size
    than the diagonal metric."""
    from oracle import mcmc as omcmc
    from pyro_b200.infer import MCMC, NUTS
    from pyro_b200.infer.mcmc import LogisticPotential
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(0)
    n = 200
    base = torch.randn(n, 1, generator=gen)
    X = torch.cat([base + 0.1 * torch.randn(n, 1, generator=gen), base + 0.1 * torch.randn(n, 1, generator=gen),
                   torch.randn(n, 1, generator=gen)], 1)
    y = (torch.rand(n, generator=gen) < torch.sigmoid(X @ torch.tensor([1.0, -1.0, 0.5]))).double()
    out = {}
    for fm in (False, True):
        k = NUTS(potential_fn=LogisticPotential(X.to(dev), y.to(dev), 10.0), native_small=False, full_mass=fm)
        mc = MCMC(k, num_samples=120, warmup_steps=150, num_chains=8, seed=1)
        mc.run()
        out[fm] = (mc.get_samples()["beta"].cpu(), float(k._adapter.step_size.mean()), k)
    chain = omcmc.NUTSChain(omcmc.logistic_potential(X, y, 10.0), 3, seed=2)
    ref, _ = chain.run(torch.zeros(3), 200, 800)
    s, eps_dense, k = out[True]
    assert torch.allclose(s.mean(0), ref.mean(0), atol=0.3) and torch.allclose(s.std(0), ref.std(0), atol=0.2)
    assert float(torch.corrcoef(s.T)[0, 1]) < -0.95
    A = k.potential.A[0].cpu()
    cov = A @ A.T
    assert float(cov[0, 1] / (cov[0, 0] * cov[1, 1]).sqrt()) < -0.9
    assert eps_dense > 2.0 * out[False][1]


def test_full_mass_nuts_correlated_posterior(emu):
    _full_mass_case("cpu")


def _slice_nuts_case(dev):
    """``NUTS(use_multinomial_sampling=False)`` (slice-sampling tree weights, pyro/infer/mcmc/nuts.py:218-229):
    same posterior as the oracle's recursive multinomial NUTS on the 3-d logistic regression."""
    from oracle import mcmc as omcmc
    from pyro_b200.infer import MCMC, NUTS
    from pyro_b200.infer.mcmc import LogisticPotential
    torch.set_default_dtype(torch.float64)
    g = load_npz("mcmc.npz")
    X, y = torch.as_tensor(g["lr.X"]).to(dev), torch.as_tensor(g["lr.y"]).to(dev)
    k = NUTS(potential_fn=LogisticPotential(X, y, 1.0), use_multinomial_sampling=False)
    mc = MCMC(k, num_samples=150, warmup_steps=100, num_chains=8, seed=5)
    mc.run()
    assert not k._use_native and not k._use_fused_hier
    s = mc.get_samples()["beta"].cpu()
    chain = omcmc.NUTSChain(omcmc.logistic_potential(X.cpu(), y.cpu(), 1.0), 3, seed=2)
    ref, _ = chain.run(torch.zeros(3, dtype=torch.float64), 150, 600)
    assert torch.allclose(s.mean(0), ref.mean(0), atol=0.12)
    assert torch.allclose(s.std(0), ref.std(0), atol=0.08)


def test_slice_sampling_nuts_posterior(emu):
    _slice_nuts_case("cpu")


def test_lazy_exp_param_semantics(emu):
    """pyro_b200/_lazyparam.py: metadata without materialising, one autograd-connected exp on first use, and the
    param store hands it out for positive constraints (pyro/params/param_store.py:125-156 semantics kept:
    ``.unconstrained()`` and the gradient reaching the stored leaf)."""
    from torch.distributions import constraints
    from pyro_b200._lazyparam import LazyExpParam
    pyro.clear_param_store()
    s = pyro.param("s", torch.full((3, 2), 0.5), constraint=constraints.positive)
    assert isinstance(s, LazyExpParam) and s._dense is None
    assert s.shape == (3, 2) and s.dim() == 2 and s.dtype == torch.get_default_dtype() and s._dense is None
    u = s.unconstrained()
    assert u.requires_grad and u.grad_fn is None and torch.allclose(u, torch.full((3, 2), 0.5).log())
    out = (s * 2.0).sum() + s.log().sum()           # any torch function materialises exp(u), once
    assert s._dense is not None and torch.allclose(s._dense, torch.full((3, 2), 0.5))
    d = s._dense
    _ = s + 1
    assert s._dense is d
    out.backward()
    assert torch.allclose(u.grad, torch.full((3, 2), 2 * 0.5 + 1.0))
    # a distribution other than the Normal draw path densifies at the autograd boundary
    pyro.clear_param_store()
    r = pyro.param("r", torch.tensor([1.5, 2.0]), constraint=constraints.positive)
    lp = dist.Gamma(r, 1.0).log_prob(torch.tensor([0.3, 0.7])).sum()
    lp.backward()
    ro = torch.tensor([1.5, 2.0]).log().requires_grad_(True)
    torch.distributions.Gamma(ro.exp(), 1.0).log_prob(torch.tensor([0.3, 0.7])).sum().backward()
    assert torch.allclose(r.unconstrained().grad, ro.grad, atol=1e-6)


@pytest.mark.parametrize("subsample", [False, True])
def test_latent_block_step_equals_sitewise_step(emu, subsample):
    """Trace_ELBO with the latent-sites block (log-scale draw of a positive parameter, Normal prior folded into
    the draw's backward, batched value-only prior scoring) gives the loss and gradients of the site-by-site
    path; the claim is really taken (both priors) and really skipped when the prior is learnable."""
    from pyro_b200 import _native as N
    from pyro_b200.infer import Trace_ELBO
    import pyro_b200.distributions as pd
    torch.manual_seed(3)
    X = torch.randn(40, 5)
    y = (torch.rand(40) < 0.4).to(X.dtype)

    def model(X, y):
        w = pyro.sample("w", dist.Normal(X.new_zeros(5), X.new_ones(5)).to_event(1))
        b = pyro.sample("b", dist.Normal(X.new_zeros(()), X.new_full((), 10.0)))
        with pyro.plate("data", 40, subsample_size=10 if subsample else None) as idx:
            lg = (X[idx] @ w.squeeze(-2).T).T + b if w.dim() > 1 else X[idx] @ w + b
            pyro.sample("y", dist.Bernoulli(logits=lg), obs=y[idx])

    res, claimed = [], []
    orig = pd.claim_rsample_prior

    def spy(fn, value, weight):
        out = orig(fn, value, weight)
        claimed.append(out is not None)
        return out

    for latent in (True, False):
        N.LATENT_BLOCK = latent
        N.LAZY_PARAM = latent
        pd.claim_rsample_prior = spy
        try:
            pyro.clear_param_store()
            torch.manual_seed(7)
            elbo = Trace_ELBO(num_particles=6, vectorize_particles=True, max_plate_nesting=1)
            loss = elbo.loss_and_grads(model, models.logistic_guide, X, y)
            store = pyro.get_param_store()
            res.append((loss, {k: store._params[k].grad.clone() for k in ("w_loc", "w_scale", "b_loc", "b_scale")}))
        finally:
            N.LATENT_BLOCK = True
            N.LAZY_PARAM = True
            pd.claim_rsample_prior = orig
    assert claimed[:2] == [True, True] and not any(claimed[2:])
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[1][0])
    for k in res[0][1]:
        assert torch.allclose(res[0][1][k], res[1][1][k], rtol=1e-5, atol=1e-6), k
