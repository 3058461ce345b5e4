"""HMC and NUTS kernels that advance ALL local chains together on the device.

Interface mirrors pyro/infer/mcmc/hmc.py:96-150 and pyro/infer/mcmc/nuts.py:137-183 (same
constructor arguments; ``setup / sample / logging / diagnostics / cleanup`` of
pyro/infer/mcmc/mcmc_kernel.py:8-80).  ``potential_fn`` may be a ``NativePotential`` /
``TracePotential`` object (the seam SURVEY.md 8b names); a ``model`` is wrapped in a
``TracePotential``.

Semantics restate the reference transition (nuts.py:367-522, hmc.py:371-438): momentum refresh,
multinomial NUTS with the generalised U-turn criterion, max tree depth, divergence threshold 1000,
biased top-level acceptance.  The reference recursion builds one chain's tree in Python; here:

* native model, small latent dim: ``b2_nuts_small`` runs whole transitions in one kernel, one
  thread per chain (pyro_b200/csrc/nuts_core.cuh);
* otherwise: an ITERATIVE tree builder in lockstep over chains.  A chain's transition ends as soon
  as one of its subtrees is cut short (U-turn / divergence), so all still-active chains always sit
  at the same (depth, leaf) -- those are host integers -- and only ``done`` masks, directions and
  the O(depth) momentum checkpoints are per chain, held on the device.  One potential evaluation
  serves all chains; the host never reads a device value inside a transition except a periodic
  "all chains done?" poll.
"""
import ctypes
import math

import torch

from ... import _native as N
from .adaptation import WarmupAdapter
from .potential import HierNormalPotential, NativePotential, TracePotential, WhitenedPotential

_MAX_SLICED_ENERGY = 1000.0


def _popcount(n):
    return bin(n).count("1")


def _trailing_ones(n):
    c = 0
    while n & 1:
        n >>= 1
        c += 1
    return c


def _logaddexp(a, b):
    # pyro/infer/mcmc/nuts.py:15-17, safe for -inf
    m = torch.maximum(a, b)
    m_safe = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(m), m)
    return m_safe + torch.log(torch.exp(a - m_safe) + torch.exp(b - m_safe))


class HMC:
    """Fixed-length Hamiltonian Monte Carlo (hmc.py:371-438) over ``C`` chains."""

    def __init__(self, model=None, potential_fn=None, step_size=1, trajectory_length=None,
                 num_steps=None, adapt_step_size=True, adapt_mass_matrix=True, full_mass=False,
                 transforms=None, max_plate_nesting=None, jit_compile=False, jit_options=None,
                 ignore_jit_warnings=False, target_accept_prob=0.8, init_strategy=None,
                 min_stepsize=1e-10, max_stepsize=1e10, compile_model=True):
        self.compile_model = compile_model   # recognise native model classes (infer/mcmc/compile.py)
        if not ((model is None) ^ (potential_fn is None)):
            raise ValueError("Only one of `model` or `potential_fn` must be specified.")
        # dense mass matrix (pyro/infer/mcmc/adaptation.py:238-392 BlockMassMatrix with full_mass=True):
        # run unit-mass dynamics in the whitened coordinates z = A z', A = chol(adapted covariance) --
        # identical trajectories, U-turn decisions and acceptance to mass matrix M = (A A^T)^-1 on z
        # (potential.WhitenedPotential); every kernel stays on its diagonal/identity-mass path
        self.full_mass = bool(full_mass)
        self.model = model
        self.potential = potential_fn
        self.step_size = step_size
        if trajectory_length is not None:
            self.trajectory_length = trajectory_length
        elif num_steps is not None:
            self.trajectory_length = step_size * num_steps
        else:
            self.trajectory_length = 2 * math.pi  # from Stan
        self.adapt_step_size = adapt_step_size
        self.adapt_mass_matrix = adapt_mass_matrix
        self.target_accept_prob = target_accept_prob
        self.max_plate_nesting = max_plate_nesting
        self._min_stepsize = min_stepsize
        self._max_stepsize = max_stepsize
        self._direction_threshold = math.log(0.8)  # from Stan
        self._reset()

    def _reset(self):
        self._t = 0
        self._warmup_steps = None
        self._z = self._U = self._g = None
        self._adapter = None
        self._divergences = None
        self._accept_cnt = None
        self._mean_accept = None
        self._gen = None
        self.num_leapfrogs = 0  # chain-leapfrogs performed (for throughput reporting)
        self._leap_dev = None   # same, counted on the device (masked chains excluded)

    # ---- setup -----------------------------------------------------------------------------------
    def setup(self, warmup_steps, num_chains, *args, seed=0, initial_params=None, **kwargs):
        self._warmup_steps = warmup_steps
        self.C = num_chains
        if self.potential is None:
            native = None
            if getattr(self, "compile_model", True):
                from .compile import recognise
                native = recognise(self.model, args, kwargs)
            self.potential = native if native is not None else TracePotential(
                self.model, args, kwargs, num_chains, max_plate_nesting=self.max_plate_nesting)
        pot = self.potential
        if isinstance(pot, TracePotential):
            pot.C = num_chains
        if self.full_mass and not isinstance(pot, WhitenedPotential):
            if pot.dim > 4096:
                raise ValueError("full_mass=True needs a [chains, D, D] factor; D = %d is too large" % pot.dim)
            pot = self.potential = WhitenedPotential(pot)
        self.D = pot.dim
        dev, dtype = pot.device, pot.dtype
        N.require_cuda(torch.empty(0, device=dev), "MCMC kernels (model / potential data)")
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(int(seed))
        self._seed = int(seed)
        # initial points: uniform(-2, 2) in unconstrained space, retried until finite
        # (pyro/infer/mcmc/util.py:325-367, init_to_uniform)
        if initial_params is not None:
            z = initial_params.to(dev, dtype).reshape(num_chains, self.D).clone()
            U, g = pot.value_and_grad(z)
        else:
            z = pot.init_uniform(num_chains, generator=self._gen)
            U, g = pot.value_and_grad(z)
            for _ in range(100):
                bad = ~(torch.isfinite(U) & torch.isfinite(g).all(-1))
                if not bool(bad.any()):
                    break
                z = torch.where(bad[:, None], pot.init_uniform(num_chains, generator=self._gen), z)
                U, g = pot.value_and_grad(z)
            else:
                raise ValueError("Model specification seems incorrect - cannot find valid initial params.")
        self._z, self._U, self._g = z, U, g
        self._adapter = WarmupAdapter(num_chains, self.D, dtype, dev, step_size=self.step_size,
                                      adapt_step_size=self.adapt_step_size,
                                      target_accept_prob=self.target_accept_prob,
                                      adapt_mass_matrix=self.adapt_mass_matrix, full_mass=self.full_mass,
                                      mass_update_fn=self._set_whitening if self.full_mass else None)
        self._adapter.configure(warmup_steps, find_reasonable_step_size_fn=self._find_reasonable_step_size)
        if self.adapt_step_size:
            self._adapter.reset_step_size_adaptation(z)
        self._divergences = torch.zeros(num_chains, dtype=torch.int64, device=dev)
        self._accept_cnt = torch.zeros(num_chains, dtype=torch.int64, device=dev)
        self._mean_accept = torch.zeros(num_chains, dtype=torch.float64, device=dev)

    @property
    def initial_params(self):
        return self._out(self._z)

    def _out(self, z):
        """Kernel coordinates -> the model's unconstrained coordinates (identity unless full_mass)."""
        return self.potential.to_original(z) if self.full_mass else z

    def _set_whitening(self, A):
        """Adopt a new whitening factor ``A`` [C, D, D] (end of a mass-adaptation window): the current
        state is re-expressed in the new coordinates and its potential / gradient re-evaluated."""
        z_orig = self.potential.to_original(self._z)
        self.potential.set_factor(A)
        self._z = self.potential.from_original(z_orig)
        self._U, self._g = self.potential.value_and_grad(self._z)
        return self._z

    # ---- pieces --------------------------------------------------------------------------------
    def _rand(self, *shape):
        return torch.rand(*shape, generator=self._gen, device=self._z.device, dtype=self._z.dtype)

    def _randn(self, *shape):
        return torch.randn(*shape, generator=self._gen, device=self._z.device, dtype=self._z.dtype)

    def _leapfrog(self, z, r, g, eps, minv, active=None):
        """One velocity-Verlet step (pyro/ops/integrator.py:45-65) for all chains through the
        C-ABI kernels; returns (z, r, g, U, kinetic)."""
        C, D = z.shape
        dev = z.device
        dt = N._DTYPES[z.dtype]
        act = active.data_ptr() if active is not None else None
        lib = N.lib()
        N.check(lib.b2_leapfrog_half_kick_drift(z.data_ptr(), r.data_ptr(), g.data_ptr(),
                                                eps.data_ptr(), minv.data_ptr(), D, act, C, D, dt,
                                                N.stream_ptr(dev)), "b2_leapfrog_half_kick_drift")
        if isinstance(self.potential, NativePotential):
            # native potentials write the new gradient over the old one (masked chains untouched)
            U, g_new = self.potential.value_and_grad(z, active, out_grad=g)
        else:
            U, g_new = self.potential.value_and_grad(z, active)
        ke = torch.empty(C, dtype=z.dtype, device=dev)
        ws = N.workspace(dev, int(lib.b2_mcmc_workspace(C)), tag="mcmc")
        N.check(lib.b2_leapfrog_half_kick(r.data_ptr(), g_new.data_ptr(), eps.data_ptr(),
                                          minv.data_ptr(), D, act, ke.data_ptr(), C, D, dt,
                                          ws.data_ptr(), ws.numel(), N.stream_ptr(dev)),
                "b2_leapfrog_half_kick")
        self.num_leapfrogs += C
        return z, r, g_new, U, ke

    def _find_reasonable_step_size(self, z, step_size, minv):
        """hmc.py:170-229, for every chain independently (masked doubling/halving)."""
        U0, g0 = self.potential.value_and_grad(z)
        eps = step_size.clone()
        sq = minv.rsqrt()

        def trial(eps):
            r = self._randn(*z.shape) * sq
            ke0 = 0.5 * (minv * r * r).sum(-1)
            _, _, _, U1, ke1 = self._leapfrog(z.clone(), r, g0.clone(), eps, minv)
            delta = (U1 + ke1) - (U0 + ke0)
            return torch.where(self._direction_threshold < -delta, 1, -1)

        direction = trial(eps)
        scale = torch.where(direction > 0, 2.0, 0.5).to(eps.dtype)
        running = torch.ones_like(direction, dtype=torch.bool)
        for _ in range(100):
            running = running & (eps > self._min_stepsize) & (eps < self._max_stepsize)
            if not bool(running.any()):
                break
            eps = torch.where(running, eps * scale, eps)
            d_new = trial(eps)
            running = running & (d_new == direction)
        return eps.clamp(self._min_stepsize, self._max_stepsize)

    # ---- one HMC transition for all chains ------------------------------------------------------
    def sample(self, params=None):
        z0, U0, g0 = self._z, self._U, self._g
        eps = self._adapter.step_size
        minv = self._adapter.inverse_mass
        r = self._randn(*z0.shape) * minv.rsqrt()
        energy0 = U0 + 0.5 * (minv * r * r).sum(-1)
        # per-chain step counts differ only through step size; use the max and mask (chains are
        # independent, extra steps of a masked chain are discarded)
        num_steps = torch.clamp((self.trajectory_length / eps).floor(), min=1).long()
        nmax = int(num_steps.max())
        z, g = z0.clone(), g0.clone()
        U, ke = U0.clone(), None
        for i in range(nmax):
            active = (num_steps > i).to(torch.uint8)
            z, r, g_new, U_new, ke_new = self._leapfrog(z, r, g, eps, minv, active)
            a = active.bool()
            g = torch.where(a[:, None], g_new, g)
            U = torch.where(a, U_new, U)
            ke = ke_new if ke is None else torch.where(a, ke_new, ke)
        energy = U + ke
        energy = torch.where(torch.isnan(energy), torch.full_like(energy, float("inf")), energy)
        delta = energy - energy0
        accept_prob = (-delta).exp().clamp(max=1.0)
        accept = self._rand(self.C) < accept_prob
        self._z = torch.where(accept[:, None], z, z0)
        self._g = torch.where(accept[:, None], g, g0)
        self._U = torch.where(accept, U, U0)
        self._post_transition(accept_prob, accept, delta > _MAX_SLICED_ENERGY)
        return self._out(self._z)

    def _post_transition(self, accept_prob, accepted, diverging):
        self._t += 1
        if self._t > self._warmup_steps:
            n = self._t - self._warmup_steps
            self._accept_cnt += accepted.long()
            self._divergences += diverging.long()
        else:
            n = self._t
            self._adapter.step(self._t, self._z, accept_prob,   # nuts.py:519 passes the incremented t
                               z_model=self._out(self._z) if self.full_mass else None)
        self._mean_accept += (accept_prob.double() - self._mean_accept) / n

    def logging(self):
        return {"step size": "{:.2e}".format(float(self._adapter.step_size.mean())),
                "acc. prob": "{:.3f}".format(float(self._mean_accept.mean()))}

    def diagnostics(self):
        return {"divergences": self._divergences.tolist(),
                "acceptance rate": (self._mean_accept).tolist()}

    def cleanup(self):
        self._reset()


class NUTS(HMC):
    """No-U-Turn sampler (pyro/infer/mcmc/nuts.py), multinomial variant."""

    def __init__(self, model=None, potential_fn=None, step_size=1, adapt_step_size=True,
                 adapt_mass_matrix=True, full_mass=False, use_multinomial_sampling=True,
                 transforms=None, max_plate_nesting=None, jit_compile=False, jit_options=None,
                 ignore_jit_warnings=False, target_accept_prob=0.8, max_tree_depth=10,
                 init_strategy=None, native_small=True, fused_leaf=True):
        super().__init__(model, potential_fn, step_size, adapt_step_size=adapt_step_size,
                         adapt_mass_matrix=adapt_mass_matrix, full_mass=full_mass,
                         transforms=transforms, max_plate_nesting=max_plate_nesting,
                         target_accept_prob=target_accept_prob, init_strategy=init_strategy)
        # slice-sampling variant (pyro/infer/mcmc/nuts.py:218-229,470-475): leaf weight 1[dE <= e] with one
        # Exponential(1) slice variable per transition instead of exp(-dE); runs on the generic lockstep
        # tree (the whole-transition / fused-leaf kernels implement the multinomial default)
        self.use_multinomial_sampling = bool(use_multinomial_sampling)
        self._max_tree_depth = max_tree_depth
        self._native_small = native_small
        self._fused_leaf = fused_leaf    # model-class leaf kernel (b2_nuts_leaf_hier) when available
        self._rng_counter = None

    def setup(self, warmup_steps, num_chains, *args, **kwargs):
        super().setup(warmup_steps, num_chains, *args, **kwargs)
        self._use_native = (self._native_small and isinstance(self.potential, NativePotential)
                            and self.D <= N.NUTS_SMALL_MAX_D and self.use_multinomial_sampling)
        self._use_fused_hier = (not self._use_native and self._fused_leaf and self.use_multinomial_sampling
                                and isinstance(self.potential, HierNormalPotential))
        self._gsc = None
        if self._use_native or self._use_fused_hier:
            self._rng_counter = torch.zeros(num_chains, dtype=torch.int64, device=self._z.device)

    # ---- native whole-transition path ---------------------------------------------------------------
    def sample_native(self, num_transitions=1, collect=False):
        """Run ``num_transitions`` NUTS transitions per chain in ONE kernel launch.  Returns
        (samples [T, C, D] or None, accept_prob [T, C], depth [T, C], diverging [T, C],
        num_steps [T, C])."""
        C, D = self.C, self.D
        dev, dtype = self._z.device, self._z.dtype
        T = num_transitions
        samples = torch.empty(T, C, D, dtype=dtype, device=dev) if collect else None
        acc = torch.empty(T, C, dtype=dtype, device=dev)
        depth = torch.empty(T, C, dtype=torch.int32, device=dev)
        div = torch.empty(T, C, dtype=torch.int32, device=dev)
        steps = torch.empty(T, C, dtype=torch.int32, device=dev)
        eps = self._adapter.step_size.contiguous()
        minv = self._adapter.inverse_mass.contiguous()
        self._z = self._z.contiguous()
        self._g = self._g.contiguous()
        N.check(N.lib().b2_nuts_small(
            ctypes.byref(self.potential._model), self._z.data_ptr(), self._U.data_ptr(),
            self._g.data_ptr(), eps.data_ptr(), minv.data_ptr(), C, T, self._max_tree_depth,
            _MAX_SLICED_ENERGY, self._seed, self._rng_counter.data_ptr(),
            samples.data_ptr() if collect else None, acc.data_ptr(), depth.data_ptr(),
            div.data_ptr(), steps.data_ptr(), N.stream_ptr(dev)), "b2_nuts_small")
        self._steps_tensor = steps
        return samples, acc, depth, div, steps

    # ---- one NUTS transition for all chains -----------------------------------------------------------
    def sample(self, params=None):
        if self._use_native:
            _, acc, depth, div, steps = self.sample_native(1)
            self._leap_dev = steps.sum() if self._leap_dev is None else self._leap_dev + steps.sum()
            self._post_transition(acc[0], torch.ones_like(div[0], dtype=torch.bool), div[0] > 0)
            return self._z
        if self._use_fused_hier:
            return self._sample_lockstep_hier()
        return self._sample_lockstep()

    # ---- lockstep tree on the fused leaf kernel (hierarchical-Normal model class, any J) -----------------
    def _leaf_hier(self, st, leaf):
        """One new leaf for every active chain: b2_nuts_leaf_hier (two launches, no tensor ops).  The
        per-call host work is one ctypes call; everything constant over a subtree is cached in ``st``."""
        call = st.get("call")
        if call is None:
            lib = N.lib()
            dev = self._z.device
            ws = N.workspace(dev, int(lib.b2_nuts_leaf_hier_workspace(self.C, self.D - 2)), tag="mcmc")
            call = st["call"] = (lib.b2_nuts_leaf_hier, ctypes.byref(self.potential._model), ctypes.byref(st["c"]),
                                 ws.data_ptr(), ws.numel(), N.stream_ptr(dev), ws)
        fn, model, cst, wptr, wn, stream, _ = call
        idx_max = _popcount(leaf >> 1)
        even = leaf % 2 == 0
        nblk = 0 if even else _trailing_ones(leaf)
        rc = fn(model, cst, leaf, idx_max if even else -1, idx_max, nblk, wptr, wn, stream)
        if rc:
            N.check(rc, "b2_nuts_leaf_hier")

    def _lockstep_struct(self, t):
        c = N.b2_nuts_lockstep()
        for k in ("zL", "rL", "zR", "rR", "dir", "gscL", "gscR", "minv", "rsub", "zs", "rck", "sck", "eps",
                  "gsc_s", "U", "Us", "energy0", "logw_sub", "sum_accept", "num_prop", "done", "diverged",
                  "take", "num_leapfrogs", "rng_counter"):
            setattr(c, k, t[k].data_ptr())
        c.minv_chain_stride = t["minv"].stride(0)
        c.seed = self._seed
        c.max_delta_energy = _MAX_SLICED_ENERGY
        c.C = self.C
        return c

    def _tree_merge(self, t, rsum):
        """rsum += rsub for the active chains and the two whole-tree U-turn dot products [C, 2]
        (b2_nuts_tree_merge: one pass over [C, D])."""
        C, D = self.C, self.D
        dev = rsum.device
        dots = torch.zeros(C, 2, dtype=rsum.dtype, device=dev)
        lib = N.lib()
        ws = N.workspace(dev, int(lib.b2_mcmc_workspace(C)), tag="mcmc")
        N.check(lib.b2_nuts_tree_merge(t["rL"].data_ptr(), t["rR"].data_ptr(), t["minv"].data_ptr(),
                                       t["minv"].stride(0), rsum.data_ptr(), t["rsub"].data_ptr(),
                                       t["done"].data_ptr(), dots.data_ptr(), C, D, N._DTYPES[rsum.dtype],
                                       ws.data_ptr(), ws.numel(), N.stream_ptr(dev)), "b2_nuts_tree_merge")
        return dots

    def _rows_copy(self, dst, src, mask):
        """dst[c] = src[c] where mask[c] (b2_rows_copy_masked; only the selected rows move)."""
        m8 = mask.to(torch.uint8)
        N.check(N.lib().b2_rows_copy_masked(dst.data_ptr(), src.data_ptr(), m8.data_ptr(), dst.shape[0],
                                            dst.shape[1], N._DTYPES[dst.dtype], N.stream_ptr(dst.device)),
                "b2_rows_copy_masked")

    def _sample_lockstep_hier(self):
        """Same transition as ``_sample_lockstep`` (nuts.py:367-522 restated iteratively, all active
        chains sharing (depth, leaf)), but every leaf is ONE call of the fused kernel pair: the
        leapfrog with recomputed local gradients, the tree vectors, and the per-chain scalar logic
        all stay on the device; the host loop only counts leaves.  The two trajectory ends live in
        two buffer sets and the end picked by ``dir`` is advanced in place (a doubling always extends
        the trajectory), so nothing of size [C, D] is selected or merged back per depth; per end only
        the two global-coordinate gradients are kept (``gsc``), not a gradient vector."""
        C, D = self.C, self.D
        dev, dtype = self._z.device, self._z.dtype
        eps_abs = self._adapter.step_size
        minv = self._adapter.inverse_mass.contiguous()
        z0, U0 = self._z, self._U
        gsc0 = self._gsc if self._gsc is not None else self._g[:, :2].contiguous()
        ru = self._randn(C, D)
        r = ru * minv.rsqrt()
        energy0 = U0 + 0.5 * (ru * ru).sum(-1)
        zp, gp, Up = z0.clone(), gsc0, U0
        rsum = ru                                   # the tree's momentum sum starts as the initial leaf
        maxd = self._max_tree_depth
        u8 = dict(dtype=torch.uint8, device=dev)
        t = {
            "zL": z0.clone(), "rL": r.clone(), "zR": z0.clone(), "rR": r, "gscL": gsc0.clone(),
            "gscR": gsc0.clone(), "dir": torch.zeros(C, **u8),
            "minv": minv, "energy0": energy0.contiguous(),
            "rck": torch.empty(maxd + 1, C, D, dtype=dtype, device=dev),
            "sck": torch.empty(maxd + 1, C, D, dtype=dtype, device=dev),
            "rsub": torch.empty(C, D, dtype=dtype, device=dev),
            "sum_accept": torch.zeros(C, dtype=dtype, device=dev),
            "num_prop": torch.zeros(C, dtype=dtype, device=dev),
            "done": torch.zeros(C, **u8), "diverged": torch.zeros(C, **u8), "take": torch.zeros(C, **u8),
            "num_leapfrogs": torch.zeros(C, dtype=torch.int32, device=dev),
            "rng_counter": self._rng_counter,
            "gsc_s": torch.empty(C, 2, dtype=dtype, device=dev),
            "U": torch.empty(C, dtype=dtype, device=dev), "Us": torch.empty(C, dtype=dtype, device=dev),
            "zs": torch.empty(C, D, dtype=dtype, device=dev),
            "logw_sub": torch.empty(C, dtype=dtype, device=dev),
            "eps": torch.empty(C, dtype=dtype, device=dev),
        }
        st = {"c": self._lockstep_struct(t), "t": t}
        logw_tree = torch.zeros(C, dtype=dtype, device=dev)
        accepted = torch.zeros(C, dtype=torch.bool, device=dev)
        depth_reached = torch.zeros(C, dtype=torch.int32, device=dev)
        done = t["done"]
        for depth in range(maxd):
            if depth > 0 and bool(done.all()):
                break
            go_right = self._rand(C) < 0.5
            t["dir"].copy_(go_right)
            t["eps"].copy_(torch.where(go_right, eps_abs, -eps_abs))
            t["logw_sub"].fill_(float("-inf"))
            t["take"].zero_()
            nleaves = 1 << depth
            for leaf in range(nleaves):
                self._leaf_hier(st, leaf)
                if (leaf & 15) == 15 and leaf + 1 < nleaves and bool(done.all()):
                    break
            # ---- merge the finished subtree (chains cut short are already `done`) ---------------
            active = ~done.bool()
            logw_sub = t["logw_sub"]
            acc_tree = active & (self._rand(C) < torch.exp(logw_sub - logw_tree))
            accepted = accepted | acc_tree
            # proposal hand-over; the last drawn leaf of a chain has not been copied into zs yet (the
            # copy rides on the NEXT leaf), so those chains take it straight from the end that grew
            last = t["take"].bool() & acc_tree
            self._rows_copy(zp, t["zs"], acc_tree & ~last)
            self._rows_copy(zp, t["zR"], last & go_right)
            self._rows_copy(zp, t["zL"], last & ~go_right)
            gp = torch.where(acc_tree[:, None], t["gsc_s"], gp)
            Up = torch.where(acc_tree, t["Us"], Up)
            depth_reached = depth_reached + active.to(torch.int32)
            dots = self._tree_merge(t, rsum)
            turning_top = (dots <= 0).any(-1)
            logw_tree = torch.where(active & ~turning_top, _logaddexp(logw_tree, logw_sub), logw_tree)
            done |= (active & turning_top).to(torch.uint8)
        self._z, self._gsc, self._U = zp, gp, Up
        self._g = None
        nl = t["num_leapfrogs"].sum()
        self._leap_dev = nl if self._leap_dev is None else self._leap_dev + nl
        accept_prob = t["sum_accept"] / t["num_prop"].clamp(min=1)
        self._last_depth = depth_reached
        self._post_transition(accept_prob, accepted, t["diverged"].bool())
        return self._z

    def _leaf_vector(self, z, rcur, g, minv, active8, take8, rsub, zs, gs, rck, sck, leaf):
        """Everything a new leaf needs over the [C, D] state in one fused pass
        (``b2_nuts_leaf_vector``).  Returns the U-turn flags [C] of the blocks ending at this leaf
        (all False on even leaves)."""
        C, D = z.shape
        dev = z.device
        idx_max = _popcount(leaf >> 1)
        even = leaf % 2 == 0
        nblk = 0 if even else _trailing_ones(leaf)
        dots = torch.empty(C, max(2 * nblk, 1), dtype=z.dtype, device=dev)
        lib = N.lib()
        ws = N.workspace(dev, int(lib.b2_mcmc_workspace(C)), tag="mcmc")
        N.check(lib.b2_nuts_leaf_vector(
            z.data_ptr(), rcur.data_ptr(), g.data_ptr(), minv.data_ptr(), D, active8.data_ptr(),
            take8.data_ptr(), rsub.data_ptr(), zs.data_ptr(), gs.data_ptr(), rck.data_ptr(),
            sck.data_ptr(), idx_max if even else -1, idx_max, nblk, dots.data_ptr(), C, D,
            N._DTYPES[z.dtype], ws.data_ptr(), ws.numel(), N.stream_ptr(dev)), "b2_nuts_leaf_vector")
        if even:
            return torch.zeros(C, dtype=torch.bool, device=dev)
        return (dots[:, : 2 * nblk] <= 0).any(-1)

    def _sample_lockstep(self):
        C, D = self.C, self.D
        pot = self.potential
        dev, dtype = self._z.device, self._z.dtype
        eps_abs = self._adapter.step_size
        minv = self._adapter.inverse_mass.contiguous()
        s = minv.sqrt()
        z0, U0, g0 = self._z, self._U, self._g
        ru = self._randn(C, D)                      # whitened momentum r_u ~ N(0, I)
        r = ru / s                                  # r = M^{1/2} r_u
        energy0 = U0 + 0.5 * (ru * ru).sum(-1)
        neg_inf = torch.full((C,), float("-inf"), dtype=dtype, device=dev)
        zl, rl, gl, rul = z0.clone(), r.clone(), g0.clone(), ru.clone()
        zr, rr, gr, rur = z0.clone(), r.clone(), g0.clone(), ru.clone()
        zp, gp, Up = z0.clone(), g0.clone(), U0.clone()
        rsum = ru.clone()
        logw_tree = torch.zeros(C, dtype=dtype, device=dev)
        sum_accept = torch.zeros(C, dtype=dtype, device=dev)
        num_prop = torch.zeros(C, dtype=dtype, device=dev)
        done = torch.zeros(C, dtype=torch.bool, device=dev)
        accepted = torch.zeros(C, dtype=torch.bool, device=dev)
        diverged = torch.zeros(C, dtype=torch.bool, device=dev)
        maxd = self._max_tree_depth
        rck = torch.empty(maxd + 1, C, D, dtype=dtype, device=dev)
        sck = torch.empty(maxd + 1, C, D, dtype=dtype, device=dev)
        depth_reached = torch.zeros(C, dtype=torch.int32, device=dev)

        # slice variable of this transition: log_slice = -energy0 - Exponential(1)
        e_slice = None if self.use_multinomial_sampling else -torch.log1p(-self._rand(C))
        for depth in range(maxd):
            if depth > 0 and bool(done.all()):
                break
            go_right = self._rand(C) < 0.5
            gr_mask = go_right[:, None]
            z = torch.where(gr_mask, zr, zl).contiguous()
            rcur = torch.where(gr_mask, rr, rl).contiguous()
            g = torch.where(gr_mask, gr, gl).contiguous()
            eps = torch.where(go_right, eps_abs, -eps_abs).contiguous()
            rsub = torch.zeros(C, D, dtype=dtype, device=dev)
            logw_sub = neg_inf.clone()
            zs, gs, Us = z.clone(), g.clone(), U0.clone()
            nleaves = 1 << depth
            for leaf in range(nleaves):
                active = ~done
                act8 = active.to(torch.uint8)
                z, rcur, g_new, U_new, ke = self._leapfrog(z, rcur, g, eps, minv, act8)
                self.num_leapfrogs -= C  # recount only active chains below
                self._leap_dev = active.sum() if self._leap_dev is None else self._leap_dev + active.sum()
                if g_new is not g:
                    g = torch.where(active[:, None], g_new, g)
                energy = U_new + ke
                energy = torch.where(torch.isnan(energy), torch.full_like(energy, float("inf")), energy)
                delta = energy - energy0
                acc_p = (-delta).exp().clamp(max=1.0)
                sum_accept = sum_accept + torch.where(active, acc_p, torch.zeros_like(acc_p))
                num_prop = num_prop + active.to(dtype)
                if self.use_multinomial_sampling:
                    div_now = active & (delta > _MAX_SLICED_ENERGY)
                    w_leaf = -delta
                else:
                    sliced = delta - e_slice                       # energy_new + log_slice
                    div_now = active & (sliced > _MAX_SLICED_ENERGY)
                    w_leaf = torch.where(sliced <= 0, torch.zeros_like(delta), neg_inf)
                if leaf == 0:
                    nw = w_leaf
                    take = active
                else:
                    nw = _logaddexp(logw_sub, w_leaf)
                    take = active & (self._rand(C) < torch.exp(w_leaf - nw))
                logw_sub = torch.where(active, nw, logw_sub)
                Us = torch.where(take, U_new, Us)
                turn_now = self._leaf_vector(z, rcur, g, minv, act8, take.to(torch.uint8), rsub, zs, gs,
                                             rck, sck, leaf)
                turn_now = turn_now & active & ~div_now
                diverged = diverged | div_now
                done = done | div_now | turn_now
                if (leaf & 15) == 15 and leaf + 1 < nleaves and bool(done.all()):
                    break
            # ---- merge the finished subtree (chains cut short are already `done`) ---------------
            active = ~done
            am = active[:, None]
            right = am & gr_mask
            left = am & ~gr_mask
            ru_c = rcur * s
            zr = torch.where(right, z, zr); rr = torch.where(right, rcur, rr)
            gr = torch.where(right, g, gr); rur = torch.where(right, ru_c, rur)
            zl = torch.where(left, z, zl); rl = torch.where(left, rcur, rl)
            gl = torch.where(left, g, gl); rul = torch.where(left, ru_c, rul)
            depth_reached = depth_reached + active.to(torch.int32)
            acc_tree = active & (self._rand(C) < torch.exp(logw_sub - logw_tree))
            accepted = accepted | acc_tree
            at = acc_tree[:, None]
            zp = torch.where(at, zs, zp)
            gp = torch.where(at, gs, gp)
            Up = torch.where(acc_tree, Us, Up)
            rsum = rsum + torch.where(am, rsub, torch.zeros_like(rsub))
            rho = rsum - 0.5 * (rul + rur)
            turning_top = ((rul * rho).sum(-1) <= 0) | ((rur * rho).sum(-1) <= 0)
            logw_tree = torch.where(active & ~turning_top, _logaddexp(logw_tree, logw_sub), logw_tree)
            done = done | (active & turning_top)
        self._z, self._g, self._U = zp, gp, Up
        accept_prob = sum_accept / num_prop.clamp(min=1)
        self._last_depth = depth_reached
        self._post_transition(accept_prob, accepted, diverged)
        return self._out(self._z)

    def leapfrog_count(self):
        """Chain-leapfrogs performed so far (one device read)."""
        return self.num_leapfrogs + (int(self._leap_dev) if self._leap_dev is not None else 0)
