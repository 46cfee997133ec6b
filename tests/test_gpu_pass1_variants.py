"""GPU (-m gpu): the two pass-1 kernels of the split modes against each other and against the fp64 oracle.

Default: `pass1_np_kernel` - the N-half pipeline (two half accumulators, K loop cut in two, epilogue of one half under the MMAs of
the other, 27 x 8 KB ring of half-split weight pieces).  DPN_P1=ts: `pass1_ts_kernel` - the strict GEMM -> epilogue chain it
replaces (same operands in tensor memory, 9 x 24 KB ring).  The library reads the switch once per process, so every variant runs
in its own subprocess; the parity suite proper (test_gpu_f16x3.py, ...) runs on the default, this file keeps the other variant
alive and pins what is allowed to differ: the two kernels run the same contractions on the same operands, only the order in which
partial products reach an accumulator differs (G2: PE6 Wd^T before h1 W2^T), so both must meet the mode's oracle tolerance and
agree with each other to the same tolerance - values, loss terms and every gradient.
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


def _run(mode, p1, tmp_path):
    out = str(tmp_path / ("%s_%s.pt" % (mode, p1)))
    env = dict(os.environ, DPN_P1=p1)
    r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT), mode, out], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


@pytest.mark.parametrize("mode,tol", [("f16x3", 1e-4), ("bf16x3", 1e-2)])
def test_pipelined_and_chained_variants_agree(mode, tol, tmp_path):
    a, b = _run(mode, "np", tmp_path), _run(mode, "ts", tmp_path)
    assert ((a["o"] - b["o"]).abs().max() / b["o"].abs().max()).item() < tol
    for name, rep in (("np", a["rep"]), ("ts", b["rep"])):
        assert max(rep["jac_rel"], rep["terms_rel"], rep["grad_rel_max"]) < tol, (name, rep)
    rel = lambda x, y: ((x - y).abs().max() / y.abs().max().clamp_min(1e-30)).item()
    assert rel(a["terms"].double(), b["terms"].double()) < tol
    worst = max(rel(x.double(), y.double()) for x, y in zip(a["grads"], b["grads"]))
    assert worst < tol, worst
