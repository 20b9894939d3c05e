"""Per-kernel timing at the NTU batch-64 shapes (CUDA events, kernels run in isolation on the launching stream).

    python tools/kernel_bench.py [--only NAME] [--reps 5] [--layers 64,128,256]

Prints, per kernel and layer width, the mean launch time and the achieved ALGORITHMIC GB/s (tensors that must cross HBM
once / time) next to the measured HBM peak -- the per-kernel rooflines quoted in DESIGN.md / profiles/.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shiftgcn_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--layers", default="64,128,256")
ap.add_argument("--n", type=int, default=128)
args = ap.parse_args()
dev = torch.device("cuda:0")
V = 25
PEAK = 6452.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, nbytes, name):
    if args.only and args.only not in name:
        return
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        flush.zero_()                                    # evict L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:44s} {ms*1e3:9.1f} us   {gbs:8.1f} GB/s   {gbs/PEAK*100:5.1f}% of {PEAK:.0f}", flush=True)


for C in [int(c) for c in args.layers.split(",")]:
    T = {64: 300, 128: 150, 256: 75}[C]
    n = args.n
    R = n * T
    a = R * V * C * 4                                   # bytes of one activation tensor
    x = torch.randn(n, T, V, C, device=dev)
    z = torch.randn(n, T, V, C, device=dev)
    h = torch.relu(torch.randn(n, T, V, C, device=dev))
    gy = torch.randn(n, T, V, C, device=dev)
    out = torch.empty(n, T, V, C, device=dev)
    W = torch.randn(C, C, device=dev) / C ** 0.5
    mm = torch.rand(V, C, device=dev) + 0.5
    tab = lambda: torch.rand(V * C, device=dev) + 0.5
    ch = lambda: torch.rand(C, device=dev) + 0.5
    ypos = torch.rand(C, device=dev) * 2 - 1
    bias = torch.randn(C, device=dev)
    wimg = ops.weight_image(W, 1, C, C, C)
    stats = torch.zeros(2 * V * C, device=dev, dtype=torch.float64)
    cstats = torch.zeros(8 * C, device=dev, dtype=torch.float64)
    dmask = torch.zeros(V * C, device=dev, dtype=torch.float64)
    dW = torch.zeros(C, C, device=dev)
    A, B_, G_ = tab(), tab(), tab()
    sc, sh, mean, inv, k1, m1, m2 = ch(), ch(), ch(), ch(), ch(), ch() * 0.01, ch() * 0.01
    tag = f"[C={C},T={T}]"
    timeit(lambda: ops.rowgemm(ops.PRO_SPATIAL, ops.EPI_ROT_RAW, in0=x, out=out, wimg=wimg, groups=R, V=V, K=C, N=C, pro_a=mm, bias=bias, stats=stats), 2 * a, "rowgemm spatial/rot_raw (2a) " + tag)
    timeit(lambda: ops.rowgemm(ops.PRO_SPATIAL, ops.EPI_ROT_FUSED, in0=x, out=out, wimg=wimg, groups=R, V=V, K=C, N=C, pro_a=mm, bias=bias, epi_a=A, epi_b=B_, res=x, relu=1), 2 * a, "rowgemm spatial/rot_fused (2a) " + tag)
    timeit(lambda: ops.bn_res_relu_fwd(z, x, out, A, B_, cstats, R * V, V, C), 3 * a, "bn_res_relu_fwd (3a) " + tag)
    timeit(lambda: ops.rowgemm(ops.PRO_LERP, ops.EPI_LINEAR, in0=h, out=out, wimg=wimg, groups=R, V=V, K=C, N=C, T=T, pro_a=sc, pro_b=sh, pro_c=ypos, bias=bias, relu=1), 2 * a, "rowgemm lerp/linear (2a) " + tag)
    timeit(lambda: ops.rowgemm(ops.PRO_LERP, ops.EPI_TSHIFT, in0=h, out=out, wimg=wimg, groups=R, V=V, K=C, N=C, T=T, pro_a=sc, pro_b=sh, pro_c=ypos, bias=bias, res2=ypos, epi_a=sc, epi_b=sh, res=x, relu=1), 3 * a, "rowgemm lerp/tshift, whole eval temporal unit (3a) " + tag)
    timeit(lambda: ops.tshift_fwd(0, q=h, ypos_eff=ypos, n_samples=n, T_in=T, T_out=T, V=V, C=C, stride=1, stats=cstats), a, "tshift_fwd stats (1a) " + tag)
    timeit(lambda: ops.tshift_fwd(1, q=h, ypos_eff=ypos, n_samples=n, T_in=T, T_out=T, V=V, C=C, stride=1, res=x, out=out, scale=sc, shift=sh, relu=1), 3 * a, "tshift_fwd apply (3a) " + tag)
    common = dict(q=h, gy=gy, y=h, relu=1, ypos_eff=ypos, mean=mean, invstd=inv, n_samples=n, T_in=T, T_out=T, V=V, C=C, stride=1)
    timeit(lambda: ops.tshift_bwd(0, sums=cstats, **common), 3 * a, "tshift_bwd stats (3a) " + tag)
    timeit(lambda: ops.tshift_bwd(1, k1=k1, m1=m1, m2=m2, dpre=out, dbias=cstats, **common), 4 * a, "tshift_bwd apply (4a) " + tag)
    pm = dict(common, y=None, relu=0)                     # g_y pre-masked by the next unit (functional._links)
    timeit(lambda: ops.tshift_bwd(0, sums=cstats, **pm), 2 * a, "tshift_bwd stats, pre-masked g_y (2a) " + tag)
    timeit(lambda: ops.tshift_bwd(1, k1=k1, m1=m1, m2=m2, dpre=out, dbias=cstats, **pm), 3 * a, "tshift_bwd apply, pre-masked g_y (3a) " + tag)
    timeit(lambda: ops.rowgemm(ops.PRO_PLAIN, ops.EPI_LINEAR, in0=gy, out=out, wimg=wimg, groups=R, V=V, K=C, N=C, relu=0), 2 * a, "rowgemm plain/linear (2a) " + tag)
    timeit(lambda: ops.wgrad(ops.WG_TEMPORAL, a_src=gy, b_src=h, b_tab0=sc, b_tab1=sh, b_tab2=ypos, dw=dW, groups=R, V=V, CA=C, CB=C, T=T), 2 * a, "wgrad temporal (2a) " + tag)
    cin = dict(dp=gy, h=h, ypos_eff=ypos, mean=mean, invstd=inv, n_samples=n, T=T, V=V, C=C)
    timeit(lambda: ops.tshift_in_bwd(0, scale=sc, shift=sh, sums=cstats, **cin), 2 * a, "tshift_in_bwd stats (2a) " + tag)
    pos = torch.zeros(C, device=dev, dtype=torch.float64)
    timeit(lambda: ops.tshift_in_bwd(1, k1=k1, m1=m1, m2=m2, gh=out, relu_h=1, z=z, zmean=A, zinvstd=B_, vd_sums=stats, scale=sc, shift=sh, pos_sums=pos, **cin), 4 * a, "tshift_in_bwd apply (4a) " + tag)
    timeit(lambda: ops.rowgemm(ops.PRO_DY, ops.EPI_SPATIAL_BWD, in0=gy, in1=z, out=out, wimg=wimg, groups=R, V=V, K=C, N=C, pro_a=A, pro_b=B_, pro_c=G_, epi_a=mm, res=gy, res2=gy, res2m=h, xin=x, red0=dmask), 6 * a, "rowgemm dy/spatial_bwd (6a) " + tag)
    timeit(lambda: ops.rowgemm(ops.PRO_DY, ops.EPI_SPATIAL_BWD, in0=gy, in1=z, out=out, wimg=wimg, groups=R, V=V, K=C, N=C, pro_a=A, pro_b=B_, pro_c=G_, epi_a=mm, res=gy, res2=h, xin=x, red0=dmask, relu=1), 5 * a, "rowgemm dy/spatial_bwd, pre-masked (5a) " + tag)
    timeit(lambda: ops.wgrad(ops.WG_SPATIAL, a_src=x, a_tab0=mm, b_src=gy, b_src2=z, b_tab0=A, b_tab1=B_, b_tab2=G_, dw=dW, groups=R, V=V, CA=C, CB=C), 3 * a, "wgrad spatial (3a) " + tag)
    del x, z, h, gy, out
    torch.cuda.empty_cache()
