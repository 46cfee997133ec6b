"""Accuracy of every arithmetic mode against the fp64 oracle over several draws (diagnostic / profiles table)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deepphysinet_b200 import testing as T
cases = [(1, 128, 128), (1, 128, 7), (1, 256, 9), (1, 700, 700), (1, 1000, 11), (2, 2048, 5), (1, 8192, 3)]
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "f16x3", "bf16x3", "bf16"]
print("%-18s %-7s %9s %9s %9s %9s  %s" % ("case (B,N,seed)", "mode", "terms", "vals", "jac", "grad max", "worst tensor"))
for B, N, seed in cases:
    W, pts = T.random_decoder_weights(B=B, N=N, seed=seed, device="cuda")
    ref = T.oracle_reference(W, pts)
    for mode in modes:
        got = T.run_library(W, pts, mode=mode)
        names = W._fields
        grel = {n: T._rel(g, r) for n, g, r in zip(names, got["grads"], ref["grads"])}
        worst = max(grel, key=grel.get)
        jac = max(T._rel(got["jac"][..., k, :], ref["jac"][..., k, :]) for k in range(6))
        vals = max(T._rel(got["vals"][..., k], ref["vals"][..., k]) for k in range(6))
        terms = ((got["terms"].cpu() - ref["terms"]).abs() / ref["terms"].abs().clamp_min(1e-300)).max().item()
        print("%-18s %-7s %9.1e %9.1e %9.1e %9.1e  %s" % ((B, N, seed), mode, terms, vals, jac, grel[worst], worst))
