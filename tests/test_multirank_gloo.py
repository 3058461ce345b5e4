"""N>1 host logic on CPU: world_size-2 gloo, native seams emulated.  Particle-sharded SVI (one
all-reduce per step over [loss, grads]) must reproduce the single-process 8-particle reference
trajectory; chain-sharded MCMC must gather every rank's chains."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, what, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cpu_emulation
        import models
        import pyro_b200 as pyro
        from conftest import load_npz
        from pyro_b200.infer import MCMC, NUTS, SVI, Trace_ELBO
        from pyro_b200.infer.mcmc import LogisticPotential
        from pyro_b200.optim import ClippedAdam
        torch.set_default_dtype(torch.float64)
        with cpu_emulation.enabled():
            if what == "svi":
                g = load_npz("svi_logistic.npz")
                X, y = torch.as_tensor(g["X"]), torch.as_tensor(g["y"])
                eps_w, eps_b = torch.as_tensor(g["eps_w"]), torch.as_tensor(g["eps_b"])
                P = int(g["P"])
                Pl = P // world
                sl = slice(rank * Pl, (rank + 1) * Pl)
                box = {"i": 0}

                def guide(X, y):
                    with models.InjectNoise({"w": eps_w[box["i"], sl], "b": eps_b[box["i"], sl]}):
                        models.logistic_guide(X, y)

                svi = SVI(models.logistic_model, guide, ClippedAdam({"lr": 0.01}),
                          Trace_ELBO(num_particles=Pl, vectorize_particles=True, max_plate_nesting=1))
                out = []
                for i in range(eps_w.shape[0]):
                    box["i"] = i
                    loss = svi.step(X, y)
                    store = pyro.get_param_store()
                    flat = torch.cat([store[k].detach().reshape(-1) for k in ("w_loc", "w_scale", "b_loc", "b_scale")])
                    out.append((loss, flat.numpy()))
                ret[rank] = out
            elif what == "svi_data":
                # data-plate sharding: same particles on every rank, each rank scores its rows with
                # the plate's size/subsample_size rescaling; mean over ranks == full-data ELBO
                g = load_npz("svi_logistic.npz")
                X, y = torch.as_tensor(g["X"]), torch.as_tensor(g["y"])
                eps_w, eps_b = torch.as_tensor(g["eps_w"]), torch.as_tensor(g["eps_b"])
                n = X.shape[0]
                idx = torch.arange(rank * n // world, (rank + 1) * n // world)
                box = {"i": 0}

                def guide(X, y, idx, n_total):
                    with models.InjectNoise({"w": eps_w[box["i"]], "b": eps_b[box["i"]]}):
                        models.logistic_guide(X, y)

                svi = SVI(models.logistic_model_sharded, guide, ClippedAdam({"lr": 0.01}),
                          Trace_ELBO(num_particles=int(g["P"]), vectorize_particles=True, max_plate_nesting=1))
                out = []
                for i in range(eps_w.shape[0]):
                    box["i"] = i
                    loss = svi.step(X[idx], y[idx], idx, n)
                    store = pyro.get_param_store()
                    flat = torch.cat([store[k].detach().reshape(-1) for k in ("w_loc", "w_scale", "b_loc", "b_scale")])
                    out.append((loss, flat.numpy()))
                ret[rank] = out
            else:
                g = load_npz("mcmc.npz")
                X, y = torch.as_tensor(g["lr.X"]), torch.as_tensor(g["lr.y"])
                mc = MCMC(NUTS(potential_fn=LogisticPotential(X, y, 1.0), native_small=False),
                          num_samples=20, warmup_steps=20, num_chains=6, seed=7, streaming_stats=True)
                mc.run()
                s = mc.get_samples(group_by_chain=True)["beta"]
                # pooled streaming statistics (two all-reduces of per-chain Welford states) against the
                # statistics of the all-gathered samples
                st = mc.streaming_stats(pooled=True)["beta"]
                flat = s.reshape(-1, s.shape[-1])
                ok = (torch.allclose(st["mean"], flat.mean(0), atol=1e-10)
                      and torch.allclose(st["variance"], flat.var(0, unbiased=True), atol=1e-10)
                      and st["n"] == flat.shape[0])
                ret[rank] = (tuple(s.shape), s[:, -1].numpy(), mc.local_chains, bool(ok))
    finally:
        dist.destroy_process_group()


def _spawn(what):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, what, ret), nprocs=2, join=True)
    return dict(ret)


@pytest.mark.timeout(300)
def test_particle_sharded_svi_matches_reference_trajectory():
    from conftest import load_npz
    g = load_npz("svi_logistic.npz")
    ret = _spawn("svi")
    for i, (ref_loss, ref_params) in enumerate(zip(g["losses_f64"], g["params_f64"])):
        for rank in (0, 1):
            loss, params = ret[rank][i]
            assert abs(loss - ref_loss) <= 1e-8 * abs(ref_loss), (rank, i)
            assert np.allclose(params, ref_params, atol=1e-8, rtol=1e-8), (rank, i)


@pytest.mark.timeout(300)
def test_data_sharded_svi_matches_reference_trajectory():
    from conftest import load_npz
    g = load_npz("svi_logistic.npz")
    ret = _spawn("svi_data")
    for i, (ref_loss, ref_params) in enumerate(zip(g["losses_f64"], g["params_f64"])):
        for rank in (0, 1):
            loss, params = ret[rank][i]
            assert abs(loss - ref_loss) <= 1e-8 * abs(ref_loss), (rank, i)
            assert np.allclose(params, ref_params, atol=1e-8, rtol=1e-8), (rank, i)


@pytest.mark.timeout(300)
def test_chain_sharded_mcmc_gathers_all_chains():
    ret = _spawn("mcmc")
    (shape0, last0, lc0, ok0), (shape1, last1, lc1, ok1) = ret[0], ret[1]
    assert shape0 == shape1 == (6, 20, 3) and lc0 == lc1 == 3
    assert ok0 and ok1                                    # cross-rank pooled streaming statistics
    assert np.allclose(last0, last1)                      # both ranks see the same gathered chains
    assert not np.allclose(last0[:3], last0[3:])          # rank streams differ (seed + first chain id)
