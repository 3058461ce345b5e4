"""MCMC driver (mirror of pyro/infer/mcmc/api.py:453-628: ``MCMC(kernel, num_samples,
warmup_steps, num_chains, ...)``, ``run``, ``get_samples``, ``diagnostics``, ``summary``).

The reference spawns one process per chain and hands samples over a Queue
(api.py:239-351); here all chains of a rank advance together on the GPU.  Under
``torch.distributed`` the chains are sharded over ranks (rank r owns chains
``[r*C/W, (r+1)*C/W)``) with no communication during warm-up or sampling; ``get_samples`` /
``diagnostics`` all-gather the kept samples (small) at the end.
"""
import torch

from .nuts import NUTS
from .stats import effective_sample_size, split_gelman_rubin


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class MCMC:
    def __init__(self, kernel, num_samples, warmup_steps=None, initial_params=None, num_chains=1,
                 hook_fn=None, mp_context=None, disable_progbar=True, disable_validation=True,
                 transforms=None, save_params=None, seed=0, streaming_stats=None):
        self.kernel = kernel
        self.num_samples = num_samples
        self.warmup_steps = warmup_steps if warmup_steps is not None else num_samples
        self.num_chains = num_chains
        self.initial_params = initial_params
        self.hook_fn = hook_fn
        self.save_params = save_params
        # BASELINE config 4 cannot keep its samples (1024 chains x 200 x 1e6 floats = 819 GB): with
        # ``save_params`` only those sites' columns are stored (pyro/infer/mcmc/api.py:465), and the
        # per-chain running mean / M2 of EVERY site are kept on the device (Welford; the role of
        # pyro/infer/mcmc/api.py:653-794 StreamingMCMC + pyro/ops/streaming.py)
        self.streaming = bool(save_params) if streaming_stats is None else bool(streaming_stats)
        self._stream = None
        self.seed = seed
        self._samples = None   # local [C_local, T, D]
        self._diagnostics = None
        self.rank, self.world = _dist_info()
        if num_chains % self.world != 0:
            raise ValueError("num_chains must be divisible by the number of ranks")
        self.local_chains = num_chains // self.world

    def run(self, *args, **kwargs):
        k = self.kernel
        C = self.local_chains
        init = None
        if self.initial_params is not None:
            init = self.initial_params
        # independent streams per rank: seed + first global chain id (api.py:107 seed + chain_id)
        k.setup(self.warmup_steps, C, *args, seed=self.seed + self.rank * C, initial_params=init,
                **kwargs)
        for t in range(self.warmup_steps):
            z = k.sample()
            if self.hook_fn is not None:
                self.hook_fn(k, z, "Warmup", t)
        T = self.num_samples
        pot = k.potential
        cols = None
        if self.save_params is not None:
            idx = [torch.arange(k.D)[pot.sites[name][0]] for name in self.save_params]
            cols = torch.cat(idx).to(k._z.device)
        if self.streaming:
            self._stream = {name: None for name in pot.sites}
            self._stream_n = 0
        if (isinstance(k, NUTS) and getattr(k, "_use_native", False) and self.hook_fn is None
                and cols is None and not self.streaming):
            samples, acc, depth, div, steps = k.sample_native(T, collect=True)
            k._leap_dev = steps.sum() if getattr(k, "_leap_dev", None) is None else k._leap_dev + steps.sum()
            k._divergences += (div > 0).sum(0)
            k._mean_accept = acc.double().mean(0)
            k._t += T
            self._samples = samples.transpose(0, 1).contiguous()
        else:
            width = k.D if cols is None else cols.numel()
            out = torch.empty(T, C, width, dtype=k._z.dtype, device=k._z.device)
            for t in range(T):
                z = k.sample()
                out[t] = z if cols is None else z.index_select(-1, cols)
                if self.streaming:
                    self._stream_update(pot, z)
                if self.hook_fn is not None:
                    self.hook_fn(k, z, "Sample", t)
            self._samples = out.transpose(0, 1).contiguous()
        self._cols = cols
        self._diagnostics = k.diagnostics()
        return self

    # ---- streaming statistics ------------------------------------------------------------------------
    def _stream_update(self, pot, z):
        """Welford update (pyro/ops/welford.py:7-51 per element) of every site's constrained value,
        per chain: state (mean, M2) lives on the device, nothing is stored per sample."""
        self._stream_n += 1
        n = self._stream_n
        for name, info in pot.sites.items():
            v = pot.unpack_site(name, z[..., info[0]])
            st = self._stream[name]
            if st is None:
                self._stream[name] = [v.clone(), torch.zeros_like(v)]
                continue
            mean, m2 = st
            delta = v - mean
            mean += delta / n
            m2 += delta * (v - mean)

    def streaming_stats(self, pooled=True):
        """Running statistics of every site over the kept transitions.  ``pooled=False``: this rank's
        per-chain ``{"mean", "variance"}`` ``[C_local, *site_shape]``.  ``pooled=True``: mean and
        (unbiased) variance over ALL chains of ALL ranks, combined on the device from the per-chain
        (n, mean, M2) with two all-reduces."""
        if self._stream is None:
            raise RuntimeError("run MCMC with save_params=... or streaming_stats=True first")
        n = self._stream_n
        out = {}
        for name, (mean, m2) in self._stream.items():
            if not pooled:
                out[name] = {"mean": mean, "variance": m2 / max(n - 1, 1), "n": n}
                continue
            chains = torch.tensor(float(mean.shape[0]), dtype=mean.dtype, device=mean.device)
            msum = mean.sum(0)
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(msum)
                dist.all_reduce(chains)
            gmean = msum / chains
            ss = (m2 + n * (mean - gmean) ** 2).sum(0)
            if self.world > 1:
                dist.all_reduce(ss)
            out[name] = {"mean": gmean, "variance": ss / (chains * n - 1).clamp(min=1), "n": int(chains) * n}
        return out

    # ---- results -------------------------------------------------------------------------------------
    def _gathered(self):
        z = self._samples
        if self.world > 1:
            import torch.distributed as dist
            parts = [torch.empty_like(z) for _ in range(self.world)]
            dist.all_gather(parts, z)
            z = torch.cat(parts, dim=0)
        return z

    def get_samples(self, num_samples=None, group_by_chain=False):
        z = self._gathered()  # [C, T, D]  (or [C, T, columns of save_params])
        pot = self.kernel.potential
        if self.save_params is not None:
            sites, off = {}, 0
            for name in self.save_params:
                sl = pot.sites[name][0]
                w = len(range(*sl.indices(self.kernel.D)))
                sites[name] = pot.unpack_site(name, z[..., off:off + w])
                off += w
        else:
            sites = pot.unpack(z)
        if not group_by_chain:
            sites = {k: v.reshape((-1,) + v.shape[2:]) for k, v in sites.items()}
        return sites

    def diagnostics(self):
        samples = self.get_samples(group_by_chain=True)
        out = {}
        for name, v in samples.items():
            d = {"n_eff": effective_sample_size(v)}
            if v.size(1) >= 4:
                d["r_hat"] = split_gelman_rubin(v)
            out[name] = d
        out.update(self._diagnostics or {})
        return out

    def summary(self, prob=0.9):
        samples = self.get_samples(group_by_chain=True)
        diag = self.diagnostics()
        rows = []
        for name, v in samples.items():
            flat = v.reshape((-1,) + v.shape[2:])
            rows.append((name, flat.mean(0), flat.std(0), diag[name]["n_eff"], diag[name].get("r_hat")))
        lines = ["{:>12} {:>10} {:>10} {:>10} {:>8}".format("site", "mean", "std", "n_eff", "r_hat")]
        for name, mean, std, neff, rhat in rows:
            m, s, n = mean.reshape(-1), std.reshape(-1), neff.reshape(-1)
            r = rhat.reshape(-1) if rhat is not None else None
            for i in range(m.numel()):
                lines.append("{:>12} {:>10.3f} {:>10.3f} {:>10.1f} {:>8}".format(
                    "{}[{}]".format(name, i), float(m[i]), float(s[i]), float(n[i]),
                    "{:.3f}".format(float(r[i])) if r is not None else "-"))
        text = "\n".join(lines)
        print(text)
        return text
