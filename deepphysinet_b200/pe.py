"""Host-side sine/cosine feature map (torch).  Layout and fp32 frequency buffer follow
DeepPhysiNet/utils/position_encoding.py:11-50; the CUDA kernels regenerate the same features on
device from the same fp32 bands (see csrc/dpn_common.cuh: pe_band)."""
import torch
import torch.nn as nn


def freq_bands(n_freqs, max_freq=4.0):
    return 2.0 ** torch.linspace(0.0, max_freq, steps=n_freqs)


class SineCosPE(nn.Module):
    def __init__(self, input_dim, N_freqs=32, max_freq=4, include_input=False):
        super().__init__()
        if include_input:
            raise NotImplementedError("the hot path only uses include_input=False")
        self.input_dim = input_dim
        self.out_dim = 2 * input_dim * N_freqs
        self.register_buffer("freq_bands", freq_bands(N_freqs, max_freq), persistent=False)

    def forward(self, inputs):
        arg = inputs[..., None, :] * self.freq_bands.to(inputs.dtype)[:, None]      # [..., F, C]
        feats = torch.stack((arg.sin(), arg.cos()), dim=-2)                          # [..., F, 2, C]
        return feats.flatten(-3)
