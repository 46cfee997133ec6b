"""Accuracy at benchmark scale: every mode against the fp64 oracle evaluated ON THE GPU (PyTorch eager, float64 - the CPU oracle
needs 16 GB and 15 s for 65 536 points), per weight tensor || g_mode - g_64 || / || g_64 ||.
    python tools/largeN_agreement.py [N] [modes]            (diagnostic; the oracle is used as the checker only)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deepphysinet_b200 import functional as Fn, testing as T
from deepphysinet_b200.config import PhysicsConsts
from oracle import dpn_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "f16x3", "bf16x3"]
W, pts = T.random_decoder_weights(B=1, N=N, seed=2, device="cuda")
consts = PhysicsConsts()
names = Fn.DecoderWeights._fields
leaves = [w.detach().double().requires_grad_(True) for w in W]
Wb = {n: (l[0] if n in ("W1", "b1", "W2", "b2", "e") else l) for n, l in zip(names, leaves)}
col = lambda k: pts[k][0].double().reshape(-1, 1)
factors = dict(zip(("motion_u_factor", "motion_v_factor", "continuous_factor", "energy_factor", "vapor_factor", "gas_factor"), consts.factor))
tot, terms, vals, jac = O.place_generated(col("x"), col("y"), col("t"), col("f"), pts["coord_data"][0].double(), Wb, dx=consts.dx, dy=consts.dy,
                                          lat_size=consts.lat_size, lon_size=consts.lon_size, pred_t_span=consts.pred_t_span,
                                          with_clip=consts.with_clip, factors=factors, return_fields=True)
tot.backward()
ref_g = [l.grad for l in leaves]
ref_terms = torch.stack([a.detach() for a in terms])
del tot, terms
torch.cuda.empty_cache()
if os.environ.get("DPN_CHUNK"):                     # experiment: smaller internal passes = more local Z-side scales in f16x3
    orig = Fn._shape
    Fn._shape = lambda *a, **kw: orig(*a, **{**kw, "chunk": int(os.environ["DPN_CHUNK"])})
for mode in modes:
    got = T.run_library(W, pts, mode=mode, want_fields=True)
    rel = {n: T._rel(g, r) for n, g, r in zip(names, got["grads"], ref_g)}
    worst = max(rel, key=rel.get)
    te = ((got["terms"][0].double() - ref_terms).abs() / ref_terms.abs()).max().item()
    jr = max(T._rel(got["jac"][0][..., k, :], jac[..., k, :]) for k in range(6))
    print("N=%d mode %-6s vs fp64 oracle: terms %.1e  jac %.1e  grads worst %.1e (%s) median %.1e  W1 %.1e W2 %.1e Wa %.1e Wd %.1e" %
          (N, mode, te, jr, rel[worst], worst, sorted(rel.values())[6], rel["W1"], rel["W2"], rel["Wa"], rel["Wd"]))
