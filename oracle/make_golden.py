"""Generates tests/golden/*.npz by running the UNMODIFIED reference (needs /root/reference).

    python -m oracle.make_golden            # from the repo root, in the build container

The reference holds no golden vectors of its own, so these files are the
pin for oracle/dpn_oracle.py: each case stores what `InterfacePhysics.place_one_batch`
(interface/interface_physics.py:271-320) + `backward()` produce in fp64 (reference made fp64-capable
by oracle/ref_harness.fp64_mode, no reference file edited) and in its native fp32, for weights
created by the reference constructors under torch.manual_seed(seed) and inputs from
dpn_oracle.synthetic_points.  Large gradient tensors are stored as (norm, strided sample).
"""
import os
import sys

import numpy as np
import torch

from . import dpn_oracle as O
from . import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
STRIDE = 997

CASES = {
    # name: dict(N, seed, geometry, with_clip, out_scale)
    "inter_0p25_n192": dict(N=192, seed=0, img=(145, 257), dx=27000.0, with_clip=True, out_scale=1.0),
    "inter_1deg_n64": dict(N=64, seed=1, img=(37, 65), dx=108000.0, with_clip=True, out_scale=1.0),
    "calibrated_n160": dict(N=160, seed=2, img=(145, 257), dx=27000.0, with_clip=True, out_scale=0.01),
    "noclip_n96": dict(N=96, seed=3, img=(145, 257), dx=27000.0, with_clip=False, out_scale=0.01),
    "single_point": dict(N=1, seed=4, img=(145, 257), dx=27000.0, with_clip=True, out_scale=0.01),
}


def make_inputs(case, dtype):
    gen = torch.Generator().manual_seed(1000 + case["seed"])
    H, Wd = case["img"]
    x, y, t, f, cd = O.synthetic_points(case["N"], gen, dx=case["dx"], dy=case["dx"], lat_size=H, lon_size=Wd,
                                        deg_per_cell=0.25 if case["dx"] < 50000 else 1.0, dtype=dtype)
    field = torch.randn(1, 159, 2405, generator=gen, dtype=torch.float64).to(dtype)
    fh = torch.tensor([[[24.0 / 360.0]]], dtype=dtype)
    return x, y, t, f, cd, field, fh


def scale_out_fc(physics_net, s):
    """'calibrated' variant of SURVEY 8(d): shrink out_fc so outputs stay near ref_data and the clip is mostly inactive."""
    if s != 1.0:
        with torch.no_grad():
            for n in O.NET_NAMES:
                getattr(physics_net, n).out_fc.weight.mul_(s)
                getattr(physics_net, n).out_fc.bias.mul_(s)


def run_reference(case, dtype):
    H, Wd = case["img"]
    ctx = rh.fp64_mode() if dtype == torch.float64 else None
    if ctx:
        ctx.__enter__()
    try:
        m, builder_loss, cfg = rh.build_reference_model(seed=case["seed"], dtype=dtype, dx=case["dx"], dy=case["dx"],
                                                        img_size=(H, Wd), with_clip=case["with_clip"])
        scale_out_fc(m.physics_net, case["out_scale"])
        x, y, t, f, cd, field, fh = make_inputs(case, dtype)
        xr, yr, tr = [a.clone().requires_grad_(True) for a in (x, y, t)]
        crit = builder_loss(name="MSELoss")
        lf = cfg["train_cfg"]["losses"]["loss_factor"]
        # the six terms exactly as place_one_batch computes them (:278-299)
        enc = m.encoding_coord(xr, yr, tr, m.pred_t_span)
        outs = m.physics_net(field, enc, cd, fh)
        u, v, P, T, q, rio = m.inverse_norm(*outs, obs_norm_cfg=m.obs_norm_cfg)
        terms = [m.montion_equation_u(xr, yr, tr, u, v, P, rio, f, crit, factor=lf["motion_u_factor"]),
                 m.montion_equation_v(xr, yr, tr, u, v, P, rio, f, crit, factor=lf["motion_v_factor"]),
                 m.continuous_equation(xr, yr, tr, u, v, rio, crit, factor=lf["continuous_factor"]),
                 m.energy_equation(xr, yr, tr, u, v, P, T, rio, q, crit, factor=lf["energy_factor"]),
                 m.vapor_equation(xr, yr, tr, u, v, P, T, q, crit, factor=lf["vapor_factor"]),
                 m.gas_equation(P, T, rio, q, crit, factor=lf["gas_factor"])]
        phys = (u, v, P, T, q, rio)
        vals = torch.cat([a.detach() for a in phys], 1)
        jac = torch.stack([torch.cat([torch.autograd.grad(a.sum(), c, retain_graph=True)[0] for c in (xr, yr, tr)], 1)
                           for a in phys], 1)
        # and the real entry point, for the total and the parameter gradients
        xr2, yr2, tr2 = [a.clone().requires_grad_(True) for a in (x, y, t)]
        total = m.place_one_batch(xr2, yr2, tr2, f, field, cd, fh, crit, lf, 0, 0, "cpu", None, "inter")
        m.zero_grad()
        total.backward()
        grads = {k: p.grad.detach().clone() for k, p in m.physics_net.named_parameters()}
        meta = m.physics_net.meta_net(field, fh).detach()
        state = {k: v.detach().clone() for k, v in m.physics_net.state_dict().items()}
    finally:
        if ctx:
            ctx.__exit__(None, None, None)
    return dict(total=total.detach(), terms=torch.stack([a.detach() for a in terms]), vals=vals, jac=jac.detach(),
                grads=grads, meta=meta, state=state, inputs=(x, y, t, f, cd, field, fh))


def main():
    if not rh.available():
        sys.exit("reference tree not mounted; golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        r64 = run_reference(case, torch.float64)
        r32 = run_reference(case, torch.float32)
        x, y, t, f, cd, field, fh = r64["inputs"]
        rec = dict(N=case["N"], seed=case["seed"], img=np.array(case["img"]), dx=case["dx"],
                   with_clip=case["with_clip"], out_scale=case["out_scale"],
                   x=x.numpy(), y=y.numpy(), t=t.numpy(), f=f.numpy(), coord_data=cd.numpy(),
                   field_checksum=np.array([field.sum().item(), field.abs().sum().item()]),
                   meta_checksum=np.array([r64["meta"].sum().item(), r64["meta"].abs().sum().item()]),
                   meta_sample=r64["meta"].flatten()[::STRIDE].numpy(),
                   total64=r64["total"].numpy(), terms64=r64["terms"].numpy(),
                   vals64=r64["vals"].numpy(), jac64=r64["jac"].numpy(),
                   total32=r32["total"].numpy(), terms32=r32["terms"].numpy(),
                   torch_version=torch.__version__)
        gnames, gnorm64, gnorm32, gdiff = [], [], [], []
        for k, g in r64["grads"].items():
            gnames.append(k)
            gnorm64.append(g.norm().item())
            gnorm32.append(r32["grads"][k].double().norm().item())
            gdiff.append((r32["grads"][k].double() - g).norm().item())
            key = "g64/" + k
            rec[key] = g.numpy() if g.numel() <= 4096 else g.flatten()[::STRIDE].numpy()
        rec["grad_names"] = np.array(gnames)
        rec["grad_norm64"] = np.array(gnorm64)
        rec["grad_norm32"] = np.array(gnorm32)
        rec["grad_ref32_vs_ref64"] = np.array(gdiff)       # the reference's own fp32 noise floor, per tensor
        rec["param_checksum"] = np.array([[v.double().sum().item(), v.double().abs().sum().item()]
                                          for v in r64["state"].values()])
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        rel32 = abs(r32["total"].item() - r64["total"].item()) / abs(r64["total"].item())
        print("%-18s total64=%.9e  ref fp32-vs-fp64 rel=%.2e  -> %s (%.0f KB)" %
              (name, r64["total"].item(), rel32, path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
