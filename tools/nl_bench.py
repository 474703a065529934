"""Hot per-launch times of fdm_norm_linear (lin_tc.cu) at the cfg4 attention shapes, next to the separate launches it replaces
(fdm_temporal_gn / fdm_gn_apply + fdm_conv 1x1).  python tools/nl_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from improved_diffusion import _native as N_  # noqa: E402
import test_gpu_kernels as K  # noqa: E402

st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


for B, T, HW in [(8, 20, 256), (8, 20, 64), (8, 20, 16)]:
    C, eps = 128, 1e-5
    M = B * T * HW
    x, gamma, beta, w, bias, g = K._nl_inputs(B, T, HW, 384, 1)
    wq = w.to(torch.bfloat16).contiguous()
    wp = wq[:128].contiguous()
    stats = K._frame_stats(x)
    tstats = torch.zeros(B, HW, 32, 2, device="cuda")
    qkv = torch.empty(M, 384, device="cuda", dtype=torch.bfloat16)
    h = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    y = torch.empty(M, C, device="cuda")
    ostats = torch.zeros(B * T, C, 2, device="cuda", dtype=torch.float64)
    xn = torch.empty(M, C, device="cuda")
    xn_op = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    side = int(HW ** 0.5)

    def nl(a_mode):
        a = N_.NormLinearArgs(x=x.data_ptr(), stats=stats.data_ptr(), tstats=tstats.data_ptr(), gamma=gamma.data_ptr(),
                              beta=beta.data_ptr(), w=wq.data_ptr(), bias=bias.data_ptr(), y_op=qkv.data_ptr(), B=B, T=T, HW=HW, K=C,
                              Cout=384, a_mode=a_mode, eps=eps)
        return lambda: N_.call("fdm_norm_linear", a, st)

    def conv(a0, Cout, resid, resid_norm=0):
        f32 = Cout == 128
        rn = dict(resid_norm=resid_norm, rn_T=T, rn_eps=eps, rn_tstats=tstats.data_ptr(), rn_stats=stats.data_ptr(),
                  rn_gamma=gamma.data_ptr(), rn_beta=beta.data_ptr()) if resid_norm else {}
        a = N_.ConvArgs(a0=a0.data_ptr(), w0=(wp if f32 else wq).data_ptr(), a1=None, w1=None, bias=bias.data_ptr(),
                        resid=(x if resid_norm else xn).data_ptr() if resid else None, y_f32=y.data_ptr() if f32 else None,
                        y_op=None if f32 else qkv.data_ptr(), stats=ostats.data_ptr() if f32 else None, N=B * T, Hin=side, Win=side,
                        C0=C, C1=0, Cout=Cout, ksize=1, stride=1, upsample=0, a_dtype=N_.BF16, op_dtype=N_.BF16, out_nchw=0,
                        engine=N_.CONV_TC, **rn)
        return lambda: N_.call("fdm_conv", a, st)

    tg = N_.TemporalGnArgs(x=x.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(), out_f32=xn.data_ptr(), out_op=xn_op.data_ptr(),
                           B=B, T=T, HW=HW, C=C, op_dtype=N_.BF16, eps=eps)
    ga = N_.GnApplyArgs(xa=x.data_ptr(), xb=None, stats_a=stats.data_ptr(), stats_b=None, gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                        film=None, out_op=xn_op.data_ptr(), out_f32=xn.data_ptr(), raw_op=None, N=B * T, HW=HW, Ca=C, Cb=0, T=T,
                        film_stride=0, film_off=0, silu=0, op_dtype=N_.BF16, eps=eps, xa_bf16=0, film_add=0)
    r = {
        "tgn": timed(lambda: N_.call("fdm_temporal_gn", tg, st)),
        "gn_apply": timed(lambda: N_.call("fdm_gn_apply", ga, st)),
        "conv qkv": timed(conv(xn_op, 384, False)),
        "conv proj+res": timed(conv(h, 128, True)),
        "conv proj+GN(x) temporal": timed(conv(h, 128, True, 2)),
        "conv proj+GN(x) frame": timed(conv(h, 128, True, 3)),
        "nl qkv spatial(a1)": timed(nl(1)),
        "nl qkv temporal(a2)": timed(nl(2)),
    }
    print(f"B={B} T={T} HW={HW}: " + "  ".join(f"{k} {v:.1f}us" for k, v in r.items() if v is not None), flush=True)
