"""GPU (-m gpu): the scaled fp16 split mode ("f16x3") against the fp64 CPU oracle.

Every operand tile is multiplied by an exact power of two derived from rigorous L1-norm bounds of the weights, split into
fp16 hi + lo (22 mantissa bits) and contracted with three MMAs (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM); the
epilogues undo the scales exactly.  Measured against the fp64 oracle (tools/mode_accuracy.py, profiles/): 2e-6..8e-6 on loss
terms, Jacobian and weight gradients where the CUDA-core fp32 mode has 5e-7 and bf16x3 1e-3 (the residue is the tensor core's
round-toward-zero fp32 accumulation over 48 MMAs per contraction), so this tensor-core mode is held to the SAME 1e-4 bound
as the fp32 mode.  As for the fp32 mode (and for the reference's own fp32 run), a draw where a ReLU / clip / delta switch
sits within rounding distance of its threshold shows an isolated 1e-4..2e-3 outlier on one field (e.g. N=256 seed 9 and
N=8192 seed 3: fp32 and f16x3 produce the SAME outlier); the seeds below are free of such ties in fp32 arithmetic.
"""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = dict(vals=1e-5, jac=1e-4, terms=1e-4, grad=1e-4)


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


@pytest.mark.parametrize("N,seed", [(1, 1), (100, 100), (128, 7), (700, 700), (1000, 11)])
def test_f16x3_random_weights_ragged_sizes(N, seed):
    _cmp(B=1, N=N, seed=seed)


def test_f16x3_batch_of_samples():
    """Three draws of a 3-sample batch: the typical error is ~3e-6; a threshold tie (see module docstring) may lift ONE draw
    to the 1e-4..2e-3 range on one field, exactly as it does for fp32 arithmetic."""
    from deepphysinet_b200 import testing as T
    worst = []
    for seed in (21, 22, 23):
        W, pts = T.random_decoder_weights(B=3, N=300, seed=seed, device="cuda")
        rep = T.compare_with_oracle(W, pts, mode="f16x3")
        print({k: v for k, v in rep.items() if k != "grad_rel"})
        assert rep["vals_rel"] < TOL["vals"], rep
        worst.append(max(rep["jac_rel"], rep["terms_rel"], rep["grad_rel_max"]))
    worst.sort()
    assert worst[1] < 1e-4 and worst[2] < 2e-3, worst


@pytest.mark.parametrize("scale", [1e-3, 30.0])
def test_f16x3_badly_scaled_weights(scale):
    """The scaling plan must keep fp16 in range whatever the magnitude of the weights (tiny / large generated weights)."""
    from deepphysinet_b200 import functional as Fn, testing as T
    W, pts = T.random_decoder_weights(B=1, N=200, seed=4, device="cuda")
    W = Fn.DecoderWeights(*[w * scale if n in ("W1", "W2", "b1") else w for n, w in zip(Fn.DecoderWeights._fields, W)])
    rep = T.compare_with_oracle(W, pts, mode="f16x3")
    print({k: v for k, v in rep.items() if k != "grad_rel"})
    assert rep["vals_rel"] < 1e-5 and rep["jac_rel"] < 1e-4 and rep["terms_rel"] < 1e-4 and rep["grad_rel_max"] < 1e-4, rep


def test_f16x3_huge_loss_factors():
    """Seeds 1e9 times larger than usual (loss factors x 1e9): the per-point and per-chunk Z-side scales follow them."""
    from deepphysinet_b200 import testing as T
    from deepphysinet_b200.config import PhysicsConsts
    base = PhysicsConsts()
    consts = PhysicsConsts(factor=tuple(f * 1e9 for f in base.factor))
    W, pts = T.random_decoder_weights(B=1, N=200, seed=4, device="cuda")
    rep = T.compare_with_oracle(W, pts, consts=consts, mode="f16x3")
    print({k: v for k, v in rep.items() if k != "grad_rel"})
    assert rep["jac_rel"] < 1e-4 and rep["terms_rel"] < 1e-4 and rep["grad_rel_max"] < 1e-4, rep


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
