"""Shared test plumbing: rebuilds, WITHOUT the reference tree, the exact model/inputs each golden
case was generated from (same seed -> same weights, checked against the stored checksums)."""
import os
import runpy

import numpy as np
import torch

from oracle import dpn_oracle as O
from oracle import make_golden as MG

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = list(MG.CASES)

META_CFG = dict(name="TransformerNet", enc_in=2405, c_out=256, d_model=256, n_heads=8, e_layers=4, d_ff=256,
                dropout=0.5, activation="gelu", output_attention=False)
NET_CFG = dict(name="PhysicsNet", in_channels=192, hidden_channels=256, out_channels=1, token_num=159,
               learnable_token_num=256)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def build_model(case, dtype=torch.float64):
    """Package PhysicsNet under the case's seed (+ out_fc calibration), on CPU."""
    from deepphysinet_b200.physics_net import PhysicsNet
    torch.manual_seed(int(case["seed"]))
    net = PhysicsNet(META_CFG, NET_CFG).to(dtype)
    MG.scale_out_fc(net, float(case["out_scale"]))      # same order as make_golden.run_reference
    return net


def case_inputs(name, dtype=torch.float64):
    spec = MG.CASES[name]
    return MG.make_inputs(spec, dtype)


def geometry(case):
    H, W = [int(v) for v in case["img"]]
    return dict(dx=float(case["dx"]), dy=float(case["dx"]), lat_size=H, lon_size=W, pred_t_span=86400.0,
                with_clip=bool(case["with_clip"]))


def leaf_weights(net, field, fh):
    """The fused operator's inputs as fp64/fp32 CPU leaf tensors (dict in oracle layout, B squeezed)."""
    with torch.no_grad():
        W = net.decoder_weights(field, fh)
    d = {}
    for k, v in W._asdict().items():
        v = v.detach()
        if k in ("W1", "b1", "W2", "b2", "e"):
            v = v[0]
        d[k] = v.clone().requires_grad_(True)
    return d


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()
