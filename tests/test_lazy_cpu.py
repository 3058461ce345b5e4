"""pyro_b200/lazy.py on the CPU: the lazy linear predictor must be invisible to model code -- every use
other than handing it to Bernoulli(logits=...) materialises exactly ``X @ w + b`` -- and must only be
created for a contraction of a latent value with a gradient-free data matrix."""
import torch
import torch.nn.functional as F

from pyro_b200.lazy import LinearPredictorTensor, SiteValue


def _mk():
    torch.manual_seed(0)
    X = torch.randn(50, 32)
    w = torch.randn(8, 1, 32, requires_grad=True)
    b = torch.randn(8, 1, requires_grad=True)
    return X, w, b


def test_vectorised_pattern_stays_lazy_and_matches_dense():
    X, w, b = _mk()
    ws, bs = SiteValue.wrap(w * 1.0), SiteValue.wrap(b * 1.0)
    lg = ws.squeeze(-2) @ X.T + bs
    assert isinstance(lg, LinearPredictorTensor) and tuple(lg.shape) == (8, 50) and lg.dim() == 2
    ref = w.squeeze(-2) @ X.T + b
    assert torch.allclose(lg.dense(), ref)
    # any other use materialises, with autograd intact
    out = (torch.sigmoid(lg) * 2).sum()
    out.backward()
    ref2 = (torch.sigmoid(ref) * 2).sum()
    gw, gb = torch.autograd.grad(ref2, [w, b])
    assert torch.allclose(w.grad, gw) and torch.allclose(b.grad, gb)


def test_single_particle_and_linear_forms():
    X, _, _ = _mk()
    w1, b1 = SiteValue.wrap(torch.randn(32)), SiteValue.wrap(torch.randn(()))
    for lg in (X @ w1 + b1, F.linear(X, w1, None) + b1, b1 + X @ w1):
        assert isinstance(lg, LinearPredictorTensor) and tuple(lg.shape) == (50,)
        assert torch.allclose(lg.dense(), X @ w1.as_subclass(torch.Tensor) + b1.as_subclass(torch.Tensor))


def test_only_data_contractions_become_lazy():
    X, w, b = _mk()
    ws = SiteValue.wrap(w * 1.0)
    q = ws.squeeze(-2) @ torch.randn(32, 5, requires_grad=True)      # other operand carries gradients
    assert type(q) is torch.Tensor
    r = ws.squeeze(-2) @ torch.randn(32, 5).t().contiguous().t()     # not the transposed view of a row-major matrix
    assert not isinstance(r, LinearPredictorTensor) or torch.allclose(r.dense(), w.squeeze(-2) @ r.lazy.X.t())
    s = ws.sum()                                                      # ordinary ops give ordinary tensors
    assert type(s) is torch.Tensor
    v = ws.squeeze(-2)
    assert isinstance(v, SiteValue)                                   # views keep the marker


def test_second_bias_or_vector_bias_materialises():
    X, w, b = _mk()
    ws, bs = SiteValue.wrap(w * 1.0), SiteValue.wrap(b * 1.0)
    lg = ws.squeeze(-2) @ X.T + bs
    twice = lg + bs                       # a second addend: no longer one affine predictor
    assert type(twice) is torch.Tensor
    assert torch.allclose(twice, w.squeeze(-2) @ X.T + 2 * b)
    per_row = (ws.squeeze(-2) @ X.T) + torch.randn(50)    # a per-ROW offset is not a per-particle bias
    assert type(per_row) is torch.Tensor


def test_distribution_constructor_keeps_it_lazy_under_validation():
    X, w, b = _mk()
    lg = SiteValue.wrap(w * 1.0).squeeze(-2) @ X.T + SiteValue.wrap(b * 1.0)
    d = torch.distributions.Bernoulli(logits=lg, validate_args=True)
    assert isinstance(d.__dict__["logits"], LinearPredictorTensor) and tuple(d.batch_shape) == (8, 50)
    # and the distribution still works if somebody uses it the ordinary way
    y = (torch.rand(50) < 0.5).float()
    lp = d.log_prob(y)
    ref = torch.distributions.Bernoulli(logits=w.squeeze(-2) @ X.T + b).log_prob(y)
    assert torch.allclose(lp, ref, atol=1e-6)
