"""Inference algorithms on the fused kernels: SVI + Trace_ELBO / TraceMeanField_ELBO, MCMC + NUTS / HMC."""
from .elbo import ELBO  # noqa: F401
from .svi import SVI  # noqa: F401
from .trace_elbo import JitTrace_ELBO, Trace_ELBO  # noqa: F401
from .trace_mean_field_elbo import TraceMeanField_ELBO  # noqa: F401
from .mcmc import HMC, MCMC, NUTS  # noqa: F401
from .renyi_elbo import RenyiELBO  # noqa: F401
from .predictive import Predictive  # noqa: F401
from .tracegraph_elbo import TraceGraph_ELBO  # noqa: F401
