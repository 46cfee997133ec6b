cd /root/repo
timeout 600 python - <<'PY' 2>&1 | tail -60 | tee gpurun_out/r02c_acc2.txt
import torch, sys
sys.path.insert(0, ".")
from deepphysinet_b200 import testing as T, functional as Fn
names = Fn.DecoderWeights._fields
def run(W, pts, mode, chunk=0):
    orig = Fn._shape
    try:
        if chunk: Fn._shape = lambda *a, **kw: orig(*a, **{**kw, "chunk": chunk})
        return T.run_library(W, pts, mode=mode)
    finally:
        Fn._shape = orig
for (B,N,seed) in [(1,1024,47),(2,640,46)]:
    W, pts = T.random_decoder_weights(B=B, N=N, seed=seed, device="cuda")
    ref = T.oracle_reference(W, pts)
    f32 = run(W, pts, "fp32")
    for mode, chunk in (("f16x3",0),("f16x3",128),("f16x3",256),("f16x3",512),("bf16",0)):
        got = run(W, pts, mode, chunk)
        rel = {n: T._rel(g, r) for n, g, r in zip(names, got["grads"], ref["grads"])}
        jac = [T._rel(got["jac"][..., k, :], ref["jac"][..., k, :]) for k in range(6)]
        te = ((got["terms"].cpu() - ref["terms"]).abs() / ref["terms"].abs()).max().item()
        print((B,N,seed), mode, "chunk", chunk, "terms %.1e jac %s |" % (te, " ".join("%.0e" % j for j in jac)), " ".join("%s %.1e" % (k, rel[k]) for k in ("W1","W2","Wa","ba","Wd","Wb")))
    rel = {n: T._rel(g, r) for n, g, r in zip(names, f32["grads"], ref["grads"])}
    print((B,N,seed), "fp32", " ".join("%s %.1e" % (k, rel[k]) for k in ("W1","W2","Wa","ba","Wd","Wb")))
    # run-to-run determinism of the f16x3 call
    a = run(W, pts, "f16x3"); b_ = run(W, pts, "f16x3")
    print("   run-to-run:", " ".join("%s %.1e" % (n, T._rel(x, y)) for n, x, y in zip(names, a["grads"], b_["grads"])))
PY
