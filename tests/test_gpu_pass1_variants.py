"""GPU (-m gpu): the two pass-1 kernels of the split modes against each other and against the fp64 oracle.

Default: `pass1_ts_kernel` - activation tile in TENSOR MEMORY (tcgen05.mma TS form, A written with tcgen05.st, 9 x 24 KB ring).
DPN_TS=0: `pass1_kernel` - activation tile in shared memory, G4 / G5 folded into one round (FOLD).  The library reads the
switch once per process, so every variant runs in its own subprocess; the parity suite proper (test_gpu_f16x3.py, ...) runs on
the default, this file keeps the other variant alive and pins what is allowed to differ:
  * values-only calls run the same arithmetic in both -> bit-identical outputs;
  * the full call differs in ONE place (q = y W2 as its own contraction instead of um (Wa W2) + 2wo W2): both variants must
    meet the mode's oracle tolerance, and agree with each other to the same tolerance.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, torch
sys.path.insert(0, %(root)r)
from deepphysinet_b200 import functional as Fn, testing as T
mode, out = sys.argv[1], sys.argv[2]
W, pts = T.random_decoder_weights(B=1, N=1000, seed=11, device="cuda")      # a draw without threshold ties (test_gpu_f16x3.py)
o = Fn.decoder_values(None, pts["coord_data"], W, xyz=(pts["x"], pts["y"], pts["t"]), mode=mode)
rep = T.compare_with_oracle(W, pts, mode=mode)
leaves = [w.clone().requires_grad_(True) for w in W]
total, terms = Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves), mode=mode)
total.backward()
torch.save(dict(o=o.cpu(), terms=terms.cpu(), grads=[l.grad.cpu() for l in leaves],
                rep={k: float(rep[k]) for k in ("jac_rel", "terms_rel", "grad_rel_max", "vals_rel")}), out)
"""


def _run(mode, ts, tmp_path):
    out = str(tmp_path / ("%s_ts%s.pt" % (mode, ts)))
    env = dict(os.environ, DPN_TS=ts)
    r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT), mode, out], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


@pytest.mark.parametrize("mode,tol", [("f16x3", 1e-4), ("bf16x3", 1e-2)])
def test_tmem_and_smem_variants_agree(mode, tol, tmp_path):
    a, b = _run(mode, "1", tmp_path), _run(mode, "0", tmp_path)
    assert torch.equal(a["o"], b["o"]), "values-only outputs must be bit-identical"
    for name, rep in (("tmem", a["rep"]), ("smem", b["rep"])):
        assert max(rep["jac_rel"], rep["terms_rel"], rep["grad_rel_max"]) < tol, (name, rep)
    rel = lambda x, y: ((x - y).abs().max() / y.abs().max().clamp_min(1e-30)).item()
    assert rel(a["terms"].double(), b["terms"].double()) < tol
    worst = max(rel(x.double(), y.double()) for x, y in zip(a["grads"], b["grads"]))
    assert worst < tol, worst
