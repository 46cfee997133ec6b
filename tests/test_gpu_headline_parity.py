"""GPU (-m gpu): parity of the HEADLINE configuration and an unbiased seed sweep ("next round" item 1).

(a) test_headline_size_vs_fp64_oracle: B = 2 x 65 536 query points (the per-sample size BASELINE.json's configs[1] quotes), both the
    CUDA-core `fp32` mode and the default tensor-core mode `f16x3` against the fp64 oracle - oracle/dpn_oracle.py, the restatement of
    interface_physics.py:271-320 that tests/test_oracle_golden.py pins to the unmodified reference - evaluated in float64 ON THE GPU
    by PyTorch (the CPU needs 16 GB and 15 s per sample at this size).  Checked: the six loss terms per sample, the 18 Jacobian
    entries per point, and every one of the 13 weight-gradient tensors.
(b) test_seed_sweep_*: 20 draws that were NOT selected for being free of threshold ties (sizes 1..2048 incl. the two draws of
    profiles/r01f_mode_accuracy.txt on which f16x3 exceeds 1e-4), the fraction of draws under 1e-4 and the worst one are asserted
    and printed.

Tolerances, stated once (DESIGN.md section 6 carries the same numbers):
  * values (no derivative, no switch in the path): fp32 1e-6, f16x3 1e-5 - every draw.
  * loss terms / Jacobian / weight gradients, `fp32` mode: <= 1e-4 on draws without a threshold tie; a ReLU pre-activation, clip
    bound or the (Dp < 0 and q >= q_s) switch within rounding distance of its threshold flips for ANY fp32 evaluation - the
    reference's own fp32 run included - and moves one field by 1e-4..1e-2 (measured on this sweep: 17 of 20 draws under 1e-4,
    median 7e-7, worst 1.1e-2 on draw (1,300,44)).  Sweep bound: >= 75 % of draws under 1e-4, median <= 5e-6, all under 5e-2.
  * same, `f16x3` (tcgen05, fp16 hi+lo operands, fp32 accumulation that rounds toward zero: a pre-activation carries 1e-6..3e-6
    where the CUDA cores have 6e-8): with ~3 million ReLU pre-activations per 1 000 points, a handful lie inside that band and
    their mask bits flip.  Most flips are invisible; one that hits a point with a dominant seed moves the Jacobian of that net and
    the gradients of its J-side tensors (Wa, ba most) by 1e-4..1.4e-3 - measured on this sweep: 10..14 of 20 draws under 1e-4
    (which draws depends on the accumulation order of the kernel version; 14 with the final kernels of round 2,
    profiles/r02w_seed_sweep_and_headline_parity.txt), median 6e-6, worst f16x3-only outlier 1.4e-3.  Values never move (<= 3e-7); loss terms stay <= 3e-5 except where the flipped switch is the
    (Dp < 0, q >= q_s) test of the vapour term itself (draw (1,128,128): 1.1e-4).  Sweep bound: loss terms <= 3e-4 on every draw
    without an fp32 tie, >= 35 % of draws entirely under 1e-4, and per draw  err(f16x3) <= max(3e-3, 2 err(fp32)).
    This is the tensor-core tolerance north_star asks to be stated separately; the strict-1e-4 mode of this library is `fp32`.
  * headline size (B = 2 x 65 536), measured: fp32 terms 1.3e-5, gradients <= 3.9e-5; f16x3 terms 9.7e-6, gradients <= 5.6e-5;
    Jacobian 1.4e-4..4.2e-4 per variable in BOTH modes (a handful of tied points among 131 072).  Bounds: terms 1e-4, every
    gradient tensor 1e-4 (fp32) / 2e-4 (f16x3), Jacobian 2e-3.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

SWEEP = [(1, 128, 128), (2, 2048, 5), (1, 1, 1), (1, 2, 2), (1, 31, 31), (1, 100, 100), (1, 127, 41), (1, 128, 7), (1, 129, 43),
         (1, 256, 9), (1, 300, 44), (3, 300, 21), (1, 511, 45), (1, 700, 700), (1, 1000, 11), (2, 640, 46), (1, 1024, 47),
         (1, 1500, 48), (1, 2048, 49), (4, 256, 50)]


def _errors(W, pts, ref, mode):
    from deepphysinet_b200 import testing as T
    got = T.run_library(W, pts, mode=mode)
    names = W._fields
    grel = {n: T._rel(g, r) for n, g, r in zip(names, got["grads"], ref["grads"])}
    jac = max(T._rel(got["jac"][..., k, :], ref["jac"][..., k, :]) for k in range(6))
    vals = max(T._rel(got["vals"][..., k], ref["vals"][..., k]) for k in range(6))
    terms = ((got["terms"].cpu() - ref["terms"].cpu()).abs() / ref["terms"].cpu().abs().clamp_min(1e-300)).max().item()
    return dict(vals=vals, terms=terms, jac=jac, grad=max(grel.values()), grad_worst=max(grel, key=grel.get), grel=grel)


@pytest.fixture(scope="module")
def sweep_table():
    from deepphysinet_b200 import testing as T
    rows = []
    for B, N, seed in SWEEP:
        W, pts = T.random_decoder_weights(B=B, N=N, seed=seed, device="cuda")
        ref = T.oracle_reference(W, pts)
        rows.append(((B, N, seed), {m: _errors(W, pts, ref, m) for m in ("fp32", "f16x3", "f16x3a")}))
    print("\n%-16s %-6s %9s %9s %9s %9s  %s" % ("case (B,N,seed)", "mode", "vals", "terms", "jac", "grad max", "tensor"))
    for case, by_mode in rows:
        for m, e in by_mode.items():
            print("%-16s %-6s %9.1e %9.1e %9.1e %9.1e  %s" % (case, m, e["vals"], e["terms"], e["jac"], e["grad"], e["grad_worst"]))
    return rows


def _summary(rows, mode):
    worst = sorted(max(e[mode]["terms"], e[mode]["jac"], e[mode]["grad"]) for _, e in rows)
    frac = sum(w < 1e-4 for w in worst) / len(worst)
    return worst, frac


def test_seed_sweep_fp32(sweep_table):
    worst, frac = _summary(sweep_table, "fp32")
    print("fp32 : %d draws, %.0f %% under 1e-4, median %.1e, worst %.1e" % (len(worst), 100 * frac, worst[len(worst) // 2], worst[-1]))
    assert all(e["fp32"]["vals"] < 1e-6 for _, e in sweep_table)
    assert worst[len(worst) // 2] < 5e-6 and frac >= 0.75 and worst[-1] < 5e-2, (frac, worst)


def test_seed_sweep_f16x3(sweep_table):
    worst, frac = _summary(sweep_table, "f16x3")
    print("f16x3: %d draws, %.0f %% under 1e-4, median %.1e, worst %.1e" % (len(worst), 100 * frac, worst[len(worst) // 2], worst[-1]))
    assert all(e["f16x3"]["vals"] < 1e-5 for _, e in sweep_table)
    assert worst[0] < 1e-5 and frac >= 0.35, (frac, worst)
    for case, e in sweep_table:                      # f16x3-only switch flips stay under 3e-3; common ties show the same outlier
        w16 = max(e["f16x3"]["terms"], e["f16x3"]["jac"], e["f16x3"]["grad"])
        w32 = max(e["fp32"]["terms"], e["fp32"]["jac"], e["fp32"]["grad"])
        assert w16 <= max(3e-3, 2.0 * w32), (case, w16, w32)
        assert e["f16x3"]["terms"] <= max(3e-4, 2.0 * e["fp32"]["terms"]), (case, e["f16x3"]["terms"])


def test_seed_sweep_f16x3a(sweep_table):
    """DPN_MODE_F16X3A: cross-first accumulation of the mask-deciding GEMMs (16 instead of 48 round-toward-zero updates per
    pre-activation).  Values at fp32-mode level (<= 3e-7), never more flips than f16x3, same per-draw bounds."""
    worst, frac = _summary(sweep_table, "f16x3a")
    _, frac16 = _summary(sweep_table, "f16x3")
    print("f16x3a: %d draws, %.0f %% under 1e-4, median %.1e, worst %.1e" % (len(worst), 100 * frac, worst[len(worst) // 2], worst[-1]))
    assert all(e["f16x3a"]["vals"] < 3e-7 for _, e in sweep_table)
    assert worst[0] < 1e-5 and frac >= max(0.35, frac16 - 0.051), (frac, frac16, worst)
    for case, e in sweep_table:
        wa = max(e["f16x3a"]["terms"], e["f16x3a"]["jac"], e["f16x3a"]["grad"])
        w32 = max(e["fp32"]["terms"], e["fp32"]["jac"], e["fp32"]["grad"])
        assert wa <= max(3e-3, 2.0 * w32), (case, wa, w32)


def _gpu_fp64_oracle(W, pts):
    """oracle/dpn_oracle.place_generated in float64 on the GPU, sample by sample (memory: ~16 GB per 65 536 points)."""
    from deepphysinet_b200 import functional as Fn
    from deepphysinet_b200.config import PhysicsConsts
    from oracle import dpn_oracle as O
    consts = PhysicsConsts()
    names = Fn.DecoderWeights._fields
    B = W.W1.shape[0]
    leaves = [w.detach().double().requires_grad_(True) for w in W]
    factors = dict(zip(("motion_u_factor", "motion_v_factor", "continuous_factor", "energy_factor", "vapor_factor", "gas_factor"),
                       consts.factor))
    terms, jacs = [], []
    for b in range(B):
        Wb = {n: (l[b] if n in ("W1", "b1", "W2", "b2", "e") else l) for n, l in zip(names, leaves)}
        col = lambda k: pts[k][b].double().reshape(-1, 1)
        tot, tt, _, jac = O.place_generated(col("x"), col("y"), col("t"), col("f"), pts["coord_data"][b].double(), Wb, dx=consts.dx,
                                            dy=consts.dy, lat_size=consts.lat_size, lon_size=consts.lon_size,
                                            pred_t_span=consts.pred_t_span, with_clip=consts.with_clip, factors=factors,
                                            return_fields=True)
        (tot / B).backward()
        terms.append(torch.stack([a.detach() for a in tt]))
        jacs.append(jac.detach())
        del tot, tt, jac
        torch.cuda.empty_cache()
    return dict(terms=torch.stack(terms), jac=torch.stack(jacs), grads=[l.grad for l in leaves])


def test_headline_size_vs_fp64_oracle():
    from deepphysinet_b200 import functional as Fn, testing as T
    W, pts = T.random_decoder_weights(B=2, N=65536, seed=2, device="cuda")
    ref = _gpu_fp64_oracle(W, pts)
    names = Fn.DecoderWeights._fields
    # bounds: loss terms / every gradient tensor / Jacobian (per variable, relative L2 over all points of both samples)
    bounds = {"fp32": dict(terms=1e-4, grad=1e-4, jac=2e-3), "f16x3": dict(terms=1e-4, grad=2e-4, jac=2e-3),
              "f16x3a": dict(terms=1e-4, grad=1e-4, jac=2e-3)}
    for mode in ("fp32", "f16x3", "f16x3a"):
        got = T.run_library(W, pts, mode=mode, want_fields=True)
        rel = {n: T._rel(g, r) for n, g, r in zip(names, got["grads"], ref["grads"])}
        te = ((got["terms"].double() - ref["terms"]).abs() / ref["terms"].abs()).max().item()
        jr = [T._rel(got["jac"][..., k, :], ref["jac"][..., k, :]) for k in range(6)]
        print("headline B=2 x 65536, %-5s vs fp64 oracle: terms %.1e | jac per var %s | grads %s" %
              (mode, te, " ".join("%.1e" % v for v in jr), " ".join("%s %.1e" % kv for kv in rel.items())))
        b = bounds[mode]
        assert te < b["terms"], (mode, te)
        assert max(jr) < b["jac"], (mode, jr)
        assert max(rel.values()) < b["grad"], (mode, rel)
        del got
        torch.cuda.empty_cache()
