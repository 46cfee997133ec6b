"""GPU (-m gpu): the scaled fp16 split mode ("f16x3") against the fp64 CPU oracle.

Every operand tile is multiplied by an exact power of two derived from rigorous L1-norm bounds of the weights, split into
fp16 hi + lo (22 mantissa bits) and contracted with three MMAs (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM); the
epilogues undo the scales exactly.

Tolerance of this mode (the same statement as tests/test_gpu_headline_parity.py and DESIGN.md section 6): the tensor core's
fp32 accumulation rounds toward zero, so a pre-activation carries 1e-6..3e-6 where the CUDA cores have 6e-8.  Values stay
<= 1e-5 on every draw; loss terms, Jacobian and weight gradients are 2e-6..8e-6 on draws where no ReLU pre-activation lies
inside that band, and a draw where one does shows an isolated 1e-4..3e-3 outlier on one net's Jacobian / J-side gradients (which
draws do depends on the accumulation order, i.e. on the kernel version).  So this file asserts PER DRAW: values 1e-5, loss
terms 3e-4, Jacobian / gradients 3e-3, and OVER EACH GROUP of draws that most of them are entirely under 1e-4.  The strict
1e-4 mode of the library is `fp32`.
"""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = dict(vals=1e-5, jac=3e-3, terms=3e-4, grad=3e-3)      # per draw (module docstring)
TYPICAL = 1e-4                                             # draws without a mask flip


def _cmp(**kw):
    from deepphysinet_b200 import testing as T
    W, pts = T.random_decoder_weights(device="cuda", **kw)
    rep = T.compare_with_oracle(W, pts, mode="f16x3")
    print({k: v for k, v in rep.items() if k != "grad_rel"})
    assert rep["vals_rel"] < TOL["vals"], rep
    assert rep["jac_rel"] < TOL["jac"], rep
    assert rep["terms_rel"] < TOL["terms"], rep
    assert rep["grad_rel_max"] < TOL["grad"], rep
    return rep


def _worst(rep):
    return max(rep["jac_rel"], rep["terms_rel"], rep["grad_rel_max"])


def test_f16x3_random_weights_ragged_sizes():
    """Five ragged sizes: every draw inside the per-draw bounds, at least four of the five entirely under 1e-4."""
    worst = sorted(_worst(_cmp(B=1, N=N, seed=seed)) for N, seed in [(1, 1), (100, 100), (128, 7), (700, 700), (1000, 11)])
    print(worst)
    assert worst[3] < TYPICAL, worst


def test_f16x3_batch_of_samples():
    """Three draws of a 3-sample batch: the typical error is ~3e-6; a mask flip (module docstring) may lift ONE draw to the
    1e-4..3e-3 range on one field."""
    from deepphysinet_b200 import testing as T
    worst = []
    for seed in (21, 22, 23):
        W, pts = T.random_decoder_weights(B=3, N=300, seed=seed, device="cuda")
        rep = T.compare_with_oracle(W, pts, mode="f16x3")
        print({k: v for k, v in rep.items() if k != "grad_rel"})
        assert rep["vals_rel"] < TOL["vals"], rep
        worst.append(_worst(rep))
    worst.sort()
    assert worst[1] < TYPICAL and worst[2] < TOL["jac"], worst


def test_f16x3_badly_scaled_weights():
    """The scaling plan must keep fp16 in range whatever the magnitude of the weights (tiny / large generated weights): an
    overflow or a flushed tile would show as an O(1) error.  Four scalings of the generated weights, per-draw bounds on each,
    at least three entirely under 1e-4."""
    from deepphysinet_b200 import functional as Fn, testing as T
    worst = []
    for scale in (1e-3, 0.05, 4.0, 30.0):
        W, pts = T.random_decoder_weights(B=1, N=200, seed=4, device="cuda")
        W = Fn.DecoderWeights(*[w * scale if n in ("W1", "W2", "b1") else w for n, w in zip(Fn.DecoderWeights._fields, W)])
        rep = T.compare_with_oracle(W, pts, mode="f16x3")
        print(scale, {k: v for k, v in rep.items() if k != "grad_rel"})
        assert rep["vals_rel"] < TOL["vals"] and rep["jac_rel"] < TOL["jac"] and rep["terms_rel"] < TOL["terms"] and rep["grad_rel_max"] < TOL["grad"], rep
        worst.append(_worst(rep))
    worst.sort()
    assert worst[2] < TYPICAL, worst


def test_f16x3_huge_loss_factors():
    """Seeds 1e9 times larger than usual (loss factors x 1e9): the per-point and per-chunk Z-side scales follow them."""
    from deepphysinet_b200 import testing as T
    from deepphysinet_b200.config import PhysicsConsts
    base = PhysicsConsts()
    consts = PhysicsConsts(factor=tuple(f * 1e9 for f in base.factor))
    W, pts = T.random_decoder_weights(B=1, N=200, seed=4, device="cuda")
    rep = T.compare_with_oracle(W, pts, consts=consts, mode="f16x3")
    print({k: v for k, v in rep.items() if k != "grad_rel"})
    assert rep["jac_rel"] < TOL["jac"] and rep["terms_rel"] < TOL["terms"] and rep["grad_rel_max"] < TOL["grad"], rep


def test_f16x3_chunking_is_invisible():
    from deepphysinet_b200 import functional as Fn, testing as T
    W, pts = T.random_decoder_weights(B=2, N=600, seed=3, device="cuda")
    a = T.run_library(W, pts, mode="f16x3")
    orig = Fn._shape
    try:
        Fn._shape = lambda *args, **kw: orig(*args, **{**kw, "chunk": 256})
        b = T.run_library(W, pts, mode="f16x3")
    finally:
        Fn._shape = orig
    assert torch.allclose(a["terms"], b["terms"], rtol=1e-6)
    for ga, gb in zip(a["grads"], b["grads"]):
        assert H.rel(ga.cpu(), gb.cpu()) < 1e-5


def test_f16x3_weight_gradients_at_scale():
    """The tensor core accumulates round-toward-zero; a weight-gradient tile that collects many point tiles in one TMEM
    accumulator drifts (65 536 points, 64 tiles per CTA: 1e-4 against an fp64 oracle).  The library flushes every 32 tiles to the
    fp32 red.add sums; pinned here against the CUDA-core fp32 mode (itself 3e-6 on these tensors, tools/largeN_agreement.py)."""
    from deepphysinet_b200 import functional as Fn, testing as T
    W, pts = T.random_decoder_weights(B=1, N=32768, seed=2, device="cuda")
    ref = T.run_library(W, pts, mode="fp32", want_fields=False)
    got = T.run_library(W, pts, mode="f16x3", want_fields=False)
    rel = {n: H.rel(g, r) for n, g, r in zip(Fn.DecoderWeights._fields, got["grads"], ref["grads"])}
    print(rel)
    assert rel["W2"] < 3e-5 and rel["Wd"] < 3e-5, rel
    assert max(rel.values()) < 2e-4, rel


def test_f16x3_full_size_properties():
    """At the size the headline is quoted on (65 536 query points per sample) the oracle is too slow; check properties that do not
    depend on it: (i) permuting the query points changes nothing but the summation order, (ii) two point-shards normalised by the
    total count (n_norm) add up to the unsharded call, (iii) loss terms and gradients are linear in the loss factors."""
    from dataclasses import replace
    from deepphysinet_b200 import functional as Fn, testing as T
    from deepphysinet_b200.config import PhysicsConsts
    N = 65536
    W, pts = T.random_decoder_weights(B=2, N=N, seed=6, device="cuda")
    full = T.run_library(W, pts, mode="f16x3", want_fields=False)
    g = torch.Generator(device="cuda").manual_seed(1)
    perm = torch.randperm(N, generator=g, device="cuda")
    shuf = T.run_library(W, {k: v[:, perm].contiguous() for k, v in pts.items()}, mode="f16x3", want_fields=False)
    assert torch.allclose(shuf["terms"], full["terms"], rtol=1e-9)             # fp64 sums
    for n, a, b in zip(Fn.DecoderWeights._fields, shuf["grads"], full["grads"]):
        assert H.rel(a, b) < 5e-5, (n, H.rel(a, b))
    terms = torch.zeros_like(full["terms"])
    grads = [torch.zeros_like(x) for x in full["grads"]]
    for lo, hi in ((0, 40000), (40000, N)):                                     # ragged shards
        out = T.run_library(W, {k: v[:, lo:hi].contiguous() for k, v in pts.items()}, mode="f16x3", want_fields=False, n_norm=N)
        terms += out["terms"]
        for acc, go in zip(grads, out["grads"]):
            acc += go
    assert torch.allclose(terms, full["terms"], rtol=1e-9)
    for n, a, b in zip(Fn.DecoderWeights._fields, grads, full["grads"]):
        assert H.rel(a, b) < 5e-5, (n, H.rel(a, b))
    base = PhysicsConsts()
    scaled = T.run_library(W, pts, consts=replace(base, factor=tuple(4.0 * f for f in base.factor)), mode="f16x3", want_fields=False)
    assert torch.allclose(scaled["terms"], 4.0 * full["terms"], rtol=1e-12)
    for n, a, b in zip(Fn.DecoderWeights._fields, scaled["grads"], full["grads"]):
        assert H.rel(a, 4.0 * b) < 2e-6, (n, H.rel(a, 4.0 * b))                   # power-of-two factor: exact up to the red.add order
