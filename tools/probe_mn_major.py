"""One-off hardware probe: which (swizzle, layout type, LBO, SBO) make MN-major TF32 operands work in tcgen05.mma."""
import ctypes, itertools, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shiftgcn_b200 import _lib, ops
from oracle.model_ref import tf32_round

lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)

def run(K, N, a_mn, b_mn, swz, layout, lbo, sbo, kstep):
    a = torch.randn((K, 128) if a_mn else (128, K), device=dev)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev)
    d = torch.full((128, N), float("nan"), device=dev)
    rc = lib.sgcn_selftest_probe(ops._p(a), ops._p(b), ops._p(d), K, N, a_mn, b_mn, swz, layout, lbo, sbo, kstep, ops._stream())
    if rc != 0:
        return "rc=%d %s" % (rc, lib.sgcn_last_error())
    try:
        torch.cuda.synchronize()
    except Exception as e:
        return "CUDA error: %s" % e
    A = tf32_round(a).double().cpu(); B = tf32_round(b).double().cpu()
    A = A.t() if a_mn else A
    B = B if b_mn else B.t()
    want = A @ B
    got = d.double().cpu()
    err = (got - want).abs().max().item() / want.abs().max().item()
    return "err=%.3e nz=%d" % (err, int((got != 0).sum()))

print("baseline K-major x K-major:", run(64, 64, 0, 0, 0, 2, 16, 1024, 1024))
KB = 16384
for (a_mn, b_mn) in ((0, 1), (1, 0), (1, 1)):
    for K in (32, 64):
        for swz, layout in ((0, 1), (1, 2), (2, 0), (0, 2), (1, 1)):
            for lbo, sbo in ((KB, 512), (512, KB), (KB, 1024), (1024, KB), (KB, 256), (256, KB)):
                r = run(K, 64, a_mn, b_mn, swz, layout, lbo, sbo, 1024)
                flag = "  <== OK" if r.startswith("err=") and float(r.split()[0][4:]) < 1e-5 else ""
                print(f"a_mn={a_mn} b_mn={b_mn} K={K} swz={swz} layout={layout} lbo={lbo} sbo={sbo}: {r}{flag}")
