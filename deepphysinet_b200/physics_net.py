"""Drop-in `PhysicsNet` / `VariableNet` (reference: DeepPhysiNet/model/physics_net.py:17-55,
DeepPhysiNet/model/variable_net.py:27-87).

Same constructor arguments, same attribute names, same parameter creation order (so the same
seed gives the same initial weights) and therefore the same `state_dict` keys/shapes - reference
checkpoints load unchanged.  What differs is what runs: the hyper-network weight generation
(variable_net.py:57-65) and the lead-time embedding (:75-78) stay in PyTorch, everything per
query point (:67-87) goes to the CUDA library through `functional.DecoderValuesFn` /
`functional.PDEResidualFn`.  There is no CPU implementation in this package.
"""
import torch
import torch.nn as nn

from . import functional as Fn
from .config import NET_ATTRS
from .encoder import MetaNet
from .pe import SineCosPE


class ResMLP(nn.Module):
    """variable_net.py:13-24 - parameter container only; the math runs inside the fused kernels."""

    def __init__(self, in_channels):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(in_channels, in_channels), nn.ReLU(inplace=True),
                                nn.Linear(in_channels, in_channels))


class VariableNet(nn.Module):
    """One hyper-network-parameterised coordinate MLP (variable_net.py:27-47)."""

    def __init__(self, token_num, in_channels, hidden_channels):
        super().__init__()
        self.in_channels, self.hidden_channels, self.token_num = in_channels, hidden_channels, token_num
        self.coord_input_fc = nn.Linear(token_num, in_channels + 1)
        self.coord_hidden_fc = nn.Linear(token_num, hidden_channels + 1)
        self.data_input_fc = nn.Linear(in_channels, hidden_channels)
        self.fore_h_fc = nn.Linear(in_channels, hidden_channels)
        self.cat_fc1 = ResMLP(hidden_channels)
        self.out_fc = nn.Linear(hidden_channels, 1)
        self.pe = SineCosPE(6, N_freqs=in_channels // 2 // 6, include_input=False)
        self.pe_fore_h = SineCosPE(1, N_freqs=in_channels // 2, include_input=False)

    def generate(self, meta_out, fore_h):
        """variable_net.py:57-65,75-78 for a batch of encoder outputs.
        meta_out [B,L,D], fore_h [B,1,1] (or [1,1,1]) -> w1 [B,D,in], b1 [B,D], w2 [B,D,hid], b2 [B,D], e [B,hid]."""
        tok = meta_out[:, : self.token_num].transpose(1, 2)                 # [B, D, token_num]
        g1 = self.coord_input_fc(tok)
        g2 = self.coord_hidden_fc(tok)
        fh = fore_h.reshape(-1, 1).expand(meta_out.shape[0], 1)
        e = self.fore_h_fc(self.pe_fore_h(fh))
        return (g1[..., : self.in_channels], g1[..., self.in_channels],
                g2[..., : self.hidden_channels], g2[..., self.hidden_channels], e)

    def forward(self, meta_out, coord, coord_data, ref_data, fore_h):
        """Reference signature (variable_net.py:49).  Values only, differentiable w.r.t. parameters and meta_out."""
        w1, b1, w2, b2, e = self.generate(meta_out.reshape(1, *meta_out.shape[-2:]), fore_h)
        W = Fn.DecoderWeights(
            W1=w1.unsqueeze(1), b1=b1.unsqueeze(1), W2=w2.unsqueeze(1), b2=b2.unsqueeze(1), e=e.unsqueeze(1),
            Wd=self.data_input_fc.weight.unsqueeze(0), bd=self.data_input_fc.bias.unsqueeze(0),
            Wa=self.cat_fc1.fc[0].weight.unsqueeze(0), ba=self.cat_fc1.fc[0].bias.unsqueeze(0),
            Wb=self.cat_fc1.fc[2].weight.unsqueeze(0), bb=self.cat_fc1.fc[2].bias.unsqueeze(0),
            wo=self.out_fc.weight.reshape(1, -1), bo=self.out_fc.bias.reshape(1))
        return Fn.decoder_values(coord, coord_data, W, ref=ref_data)


class PhysicsNet(nn.Module):
    """physics_net.py:17-55."""

    def __init__(self, meta_cfg: dict, net_cfg: dict):
        super().__init__()
        in_channels = net_cfg["in_channels"]
        hidden = net_cfg["hidden_channels"]
        token_num = net_cfg["learnable_token_num"]
        if (in_channels, hidden) != (192, 256):
            raise NotImplementedError("the sm_100a kernels are specialised for in_channels=192, hidden_channels=256 "
                                      "(configs/DeepPhysiNet_NCEP_cfg.py:25-32); got %r" % ((in_channels, hidden),))
        self.meta_net = MetaNet(meta_cfg)
        # creation order of physics_net.py:25-30 (rio before q) - it fixes the RNG stream and state_dict order
        self.U_net = VariableNet(token_num, in_channels, hidden)
        self.V_net = VariableNet(token_num, in_channels, hidden)
        self.P_net = VariableNet(token_num, in_channels, hidden)
        self.T_net = VariableNet(token_num, in_channels, hidden)
        self.rio_net = VariableNet(token_num, in_channels, hidden)
        self.q_net = VariableNet(token_num, in_channels, hidden)
        self.tanh = nn.Tanh()

    @property
    def nets(self):
        """The six decoders in output order (u, v, p, T, q, rio) = coord_data column order (physics_net.py:49-54)."""
        return [getattr(self, n) for n in NET_ATTRS]

    def decoder_weights(self, field_x, forecast_h) -> "Fn.DecoderWeights":
        """Encoder + hyper-network (PyTorch, differentiable): everything the fused operator consumes.
        field_x [B,159,2405]; forecast_h [B,1,1] or [1,1,1].  Shapes: generated [B,6,...], static [6,...]."""
        return self.decoder_weights_from_meta(self.meta_net(field_x, forecast_h), forecast_h)

    def decoder_weights_from_meta(self, meta, forecast_h) -> "Fn.DecoderWeights":
        """The hyper-network of all six nets as ONE GEMM (variable_net.py:57-65 applies two nn.Linear per net to the same
        token matrix: 12 small GEMMs forward, 24 backward, plus 5 x 6 slices to stack).  The 12 weight matrices are
        concatenated row-wise - [6 x 192 rows of coord_input_fc (w1) | 6 x 256 rows of coord_hidden_fc (w2) | 6 b1 rows | 6 b2 rows]
        = 2 700 output columns - so one [B*256, 256] x [256, 2 700] product yields every generated tensor; the lead-time
        embedding e (:75-78) is one [B,192] x [192, 6*256] product.  Same arithmetic per output element as the per-net Linears
        (fp32 dot products over the 256 tokens), parameters untouched (state_dict unchanged).  With fixed `meta` and
        `forecast_h` (a lead-time sweep over one encoded field, interface_physics.py:538-606) the result can be cached by the
        caller: `predict_grid` does."""
        nets = self.nets
        K, C, Hd, T = len(nets), nets[0].in_channels, nets[0].hidden_channels, nets[0].token_num
        B = meta.shape[0]
        wall = torch.cat([n.coord_input_fc.weight[:C] for n in nets] + [n.coord_hidden_fc.weight[:Hd] for n in nets] +
                         [n.coord_input_fc.weight[C:] for n in nets] + [n.coord_hidden_fc.weight[Hd:] for n in nets])   # [2700, T]
        ball = torch.cat([n.coord_input_fc.bias[:C] for n in nets] + [n.coord_hidden_fc.bias[:Hd] for n in nets] +
                         [n.coord_input_fc.bias[C:] for n in nets] + [n.coord_hidden_fc.bias[Hd:] for n in nets])
        tok = meta[:, :T].transpose(1, 2)                                   # [B, D, T]: hidden index = encoder channel (A.3)
        g = torch.nn.functional.linear(tok, wall, ball)                     # [B, D, 2700]
        D = g.shape[1]
        o1, o2 = K * C, K * C + K * Hd
        W1 = g[..., :o1].reshape(B, D, K, C).permute(0, 2, 1, 3)            # [B, K, D, C]
        W2 = g[..., o1:o2].reshape(B, D, K, Hd).permute(0, 2, 1, 3)         # [B, K, D, Hd]
        b1 = g[..., o2:o2 + K].transpose(1, 2)                              # [B, K, D]
        b2 = g[..., o2 + K:].transpose(1, 2)
        fh = forecast_h.reshape(-1, 1).expand(B, 1)
        pe = nets[0].pe_fore_h(fh)                                          # identical fixed encoding in every net
        e = torch.nn.functional.linear(pe, torch.cat([n.fore_h_fc.weight for n in nets]),
                                       torch.cat([n.fore_h_fc.bias for n in nets])).reshape(B, K, Hd)
        st = lambda f: torch.stack([f(n) for n in nets])
        return Fn.DecoderWeights(
            W1=W1, b1=b1, W2=W2, b2=b2, e=e,
            Wd=st(lambda n: n.data_input_fc.weight), bd=st(lambda n: n.data_input_fc.bias),
            Wa=st(lambda n: n.cat_fc1.fc[0].weight), ba=st(lambda n: n.cat_fc1.fc[0].bias),
            Wb=st(lambda n: n.cat_fc1.fc[2].weight), bb=st(lambda n: n.cat_fc1.fc[2].bias),
            wo=st(lambda n: n.out_fc.weight.reshape(-1)), bo=st(lambda n: n.out_fc.bias.reshape(())))

    def decoder_weights_per_net(self, field_x, forecast_h) -> "Fn.DecoderWeights":
        """The same tensors through VariableNet.generate, net by net (the literal structure of variable_net.py:57-65) - kept as
        the cross-check of the batched hyper-network (tests/test_hypernet_batched.py)."""
        meta = self.meta_net(field_x, forecast_h)
        gen = [n.generate(meta, forecast_h) for n in self.nets]
        st = lambda f: torch.stack([f(n) for n in self.nets])
        return Fn.DecoderWeights(
            W1=torch.stack([g[0] for g in gen], 1), b1=torch.stack([g[1] for g in gen], 1),
            W2=torch.stack([g[2] for g in gen], 1), b2=torch.stack([g[3] for g in gen], 1),
            e=torch.stack([g[4] for g in gen], 1),
            Wd=st(lambda n: n.data_input_fc.weight), bd=st(lambda n: n.data_input_fc.bias),
            Wa=st(lambda n: n.cat_fc1.fc[0].weight), ba=st(lambda n: n.cat_fc1.fc[0].bias),
            Wb=st(lambda n: n.cat_fc1.fc[2].weight), bb=st(lambda n: n.cat_fc1.fc[2].bias),
            wo=st(lambda n: n.out_fc.weight.reshape(-1)), bo=st(lambda n: n.out_fc.bias.reshape(())))

    def forward(self, field_x, coord_x, coord_data, forecast_h):
        """Reference signature (physics_net.py:41).  coord_x is the already encoded coordinate [N,192];
        returns the six normalised outputs (U, V, P, T, q, rio), each [N,1]."""
        W = self.decoder_weights(field_x, forecast_h)
        o = Fn.decoder_values(coord_x, coord_data, W)                       # [N,6]
        return tuple(o[:, i:i + 1] for i in range(6))
