"""Micro A/B on cuda:0: SDPA (memory-efficient kernel, what the encoder uses) vs explicit softmax(QK^T)V for the encoder's shape, fp32 fwd + bwd."""
import torch, torch.nn.functional as F
B, H, L, E = 8, 8, 415, 32
q, k, v = [torch.randn(B, H, L, E, device="cuda", requires_grad=True) for _ in range(3)]
def sdpa(): return F.scaled_dot_product_attention(q, k, v)
def math():
    s = torch.softmax(q @ k.transpose(-1, -2) * (E ** -0.5), dim=-1)
    return s @ v
for name, fn in (("sdpa", sdpa), ("math", math)):
    for _ in range(3): fn().sum().backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(20): o = fn()
    e[1].record()
    for _ in range(20): fn().sum().backward()
    e[2].record(); torch.cuda.synchronize()
    f = e[0].elapsed_time(e[1]) / 20; fb = e[1].elapsed_time(e[2]) / 20
    print("%s: fwd %.3f ms, fwd+bwd %.3f ms" % (name, f, fb))
print("max |sdpa - math| =", (sdpa() - math()).abs().max().item())
