"""CPU: what each tensor-core mode's operand rounding costs, predicted by emulating it in the closed-form twin
(oracle/closed_form.py).  These are the figures DESIGN.md section 6 derives the stated tolerances from; the GPU tests
(test_gpu_bf16.py / test_gpu_bf16x3.py) check that the kernels reproduce the emulation."""
import torch

from oracle import closed_form as CF
from oracle import dpn_oracle as O


def _case(N=96, seed=5):
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, fan: (torch.rand(*s, generator=g, dtype=torch.float64) * 2 - 1) / fan ** 0.5
    W = dict(W1=u(6, 256, 192, fan=192) * 2, b1=u(6, 256, fan=192), W2=u(6, 256, 256, fan=256) * 2, b2=u(6, 256, fan=256),
             e=u(6, 256, fan=192), Wd=u(6, 256, 192, fan=192), bd=u(6, 256, fan=192), Wa=u(6, 256, 256, fan=256),
             ba=u(6, 256, fan=256), Wb=u(6, 256, 256, fan=256), bb=u(6, 256, fan=256), wo=u(6, 256, fan=256) * 0.05,
             bo=u(6, fan=256) * 0.05)
    x, y, t, f, cd = O.synthetic_points(N, g)
    return (x.double(), y.double(), t.double(), f.double(), cd.double(), W)


def _worst(got, ref):
    rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-300)).item()
    terms = ((got[0] - ref[0]).abs() / ref[0].abs()).max().item()
    jac = max(rel(got[3][:, k], ref[3][:, k]) for k in range(6))
    grad = max(rel(got[1][n], ref[1][n]) for n in ref[1])
    return terms, jac, grad


def test_operand_rounding_of_each_mode():
    args = _case()
    ref = CF.pde_fwd_bwd(*args)
    bf16 = _worst(CF.pde_fwd_bwd(*args, rnd=CF.bf16_round), ref)
    bf16x3 = _worst(CF.pde_fwd_bwd(*args, rnd=CF.bf16_split_round), ref)
    f16x3 = _worst(CF.pde_fwd_bwd(*args, rnd=CF.f16_split_round), ref)
    fp32 = _worst(CF.pde_fwd_bwd(*args, rnd=lambda a: a.float().double()), ref)
    print("terms / Jacobian / worst gradient:  bf16 %s  bf16x3 %s  f16x3 %s  fp32 operands %s" % (bf16, bf16x3, f16x3, fp32))
    assert max(bf16) < 0.2                        # the stated tolerance of the bf16 mode
    assert max(bf16x3) < 1e-2 and max(bf16x3) < 0.1 * max(bf16)
    assert max(f16x3) < 1e-5 and max(f16x3) < 20 * max(fp32)      # 22 mantissa bits: within a small factor of fp32 operands
