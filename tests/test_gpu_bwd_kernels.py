"""
GPU parity of the BACKWARD C-ABI entry points (-m gpu): each one against torch.autograd over the plain PyTorch op the reference
runs at that site, on the same device, fed the SAME (bf16-rounded where applicable) operands — so tolerances only cover
accumulation order and the rounding of bf16 outputs.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    """The torch reference ops must be true fp32 (cuDNN / cuBLAS default to TF32 for convs / allow it for matmuls)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def N_():
    from improved_diffusion import _native
    return _native


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dt(t):
    return N_().BF16 if t.dtype == torch.bfloat16 else N_().F32


def dev_array(structs):
    arr = (type(structs[0]) * len(structs))(*structs)
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()


def pack(w, mode):
    n = N_()
    co, ci = w.shape[:2]
    k = w.shape[2] if w.dim() == 4 else 1
    r16, r64 = lambda v: (v + 15) // 16 * 16, lambda v: (v + 63) // 64 * 64
    if mode == n.PACK_TC_FWD:
        dst = torch.zeros(k * k, r16(co), r64(ci), dtype=torch.bfloat16, device="cuda")
    elif mode == n.PACK_TC_DGRAD:
        dst = torch.zeros(k * k, r16(ci), r64(co), dtype=torch.bfloat16, device="cuda")
    elif mode == n.PACK_SIMT_FWD:
        dst = torch.zeros(k * k, ci, co, device="cuda")
    else:
        dst = torch.zeros(k * k, co, ci, device="cuda")
    pr = n.PackProblem(src=ptr(w), src2=None, dst=ptr(dst), co=co, ci=ci, k=k, mode=mode)
    probs = dev_array([pr])
    n.call("fdm_pack_weights", n.PackWeightsArgs(problems=ptr(probs), count=1, max_elems=w.numel()), stream())
    torch.cuda.synchronize()
    return dst


def conv_native(x, wp, Cout, k, *, engine, stride=1, want="op", Cin=None):
    n = N_()
    Nf, H, W, C0 = x.shape
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty(Nf, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16 if want == "op" and x.dtype == torch.bfloat16 else torch.float32)
    a = n.ConvArgs(a0=ptr(x), w0=ptr(wp), a1=None, w1=None, bias=None, resid=None,
                   y_f32=ptr(y) if y.dtype == torch.float32 else None, y_op=ptr(y) if y.dtype == torch.bfloat16 else None,
                   stats=None, N=Nf, Hin=H, Win=W, C0=C0, C1=0, Cout=Cout, ksize=k, stride=stride, upsample=0,
                   a_dtype=dt(x), op_dtype=dt(x), out_nchw=0, engine=engine)
    n.call("fdm_conv", a, stream())
    torch.cuda.synchronize()
    return y


def test_pack_weights_all_modes():
    n = N_()
    torch.manual_seed(0)
    for co, ci, k in [(64, 64, 3), (96, 40, 3), (4, 64, 3), (64, 5, 3), (192, 64, 1)]:
        w = torch.randn(co, ci, k, k, device="cuda")
        r16, r64 = lambda v: (v + 15) // 16 * 16, lambda v: (v + 63) // 64 * 64
        got = pack(w, n.PACK_TC_FWD)
        ref = torch.zeros_like(got)
        ref[:, :co, :ci] = w.permute(3, 2, 0, 1).reshape(k * k, co, ci).to(torch.bfloat16)
        assert torch.equal(got, ref)
        wd = w.flip(2, 3).transpose(0, 1).contiguous()  # dgrad conv weights [ci][co][r][s]
        got = pack(w, n.PACK_TC_DGRAD)
        ref = torch.zeros_like(got)
        ref[:, :ci, :co] = wd.permute(3, 2, 0, 1).reshape(k * k, ci, co).to(torch.bfloat16)
        assert torch.equal(got, ref)
        assert torch.equal(pack(w, n.PACK_SIMT_FWD), w.permute(2, 3, 1, 0).reshape(k * k, ci, co))
        assert torch.equal(pack(w, n.PACK_SIMT_DGRAD), wd.permute(2, 3, 1, 0).reshape(k * k, co, ci))


@pytest.mark.parametrize("engine", ["tc", "simt"])
@pytest.mark.parametrize("case", [(5, 32, 32, 64, 64, 3), (4, 16, 16, 128, 64, 3), (6, 8, 8, 64, 192, 1), (3, 32, 32, 64, 4, 3)])
def test_conv_dgrad_through_fdm_conv(engine, case):
    """dX of a stride-1 conv = fdm_conv over the 180-degree-rotated, channel-swapped weights (FDM_PACK_*_DGRAD)."""
    n = N_()
    Nf, H, W, Ci, Co, k = case
    torch.manual_seed(1)
    w = (torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5)
    Cop = (Co + 7) // 8 * 8  # the gradient operand is channel-padded to 8 for tcgen05 (head conv: 4 -> 8)
    dy = torch.zeros(Nf, H, W, Cop, device="cuda")
    dy[..., :Co] = torch.randn(Nf, H, W, Co, device="cuda")
    if engine == "tc":
        dyb, wb = dy.to(torch.bfloat16), w.to(torch.bfloat16).float()
        wp = pack(w, n.PACK_TC_DGRAD)
        got = conv_native(dyb, wp, Ci, k, engine=n.CONV_TC).float()
        ref = torch.nn.grad.conv2d_input((Nf, Ci, H, W), wb, dyb.float()[..., :Co].permute(0, 3, 1, 2), padding=k // 2)
        tol = 6e-3  # bf16 output rounding
    else:
        wp = pack(w, n.PACK_SIMT_DGRAD)
        got = conv_native(dy[..., :Co].contiguous(), wp, Ci, k, engine=n.CONV_SIMT, want="f32")
        ref = torch.nn.grad.conv2d_input((Nf, Ci, H, W), w, dy[..., :Co].permute(0, 3, 1, 2), padding=k // 2)
        tol = 1e-5
    assert rel(got, ref.permute(0, 2, 3, 1)) <= tol


@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_stride2_dgrad_via_zero_insertion(engine):
    n = N_()
    Nf, H, W, Cc = 4, 32, 32, 64
    torch.manual_seed(2)
    w = torch.randn(Cc, Cc, 3, 3, device="cuda") / (Cc * 9) ** 0.5
    dy = torch.randn(Nf, H // 2, W // 2, Cc, device="cuda")
    z = torch.empty(Nf, H, W, Cc, device="cuda", dtype=torch.bfloat16 if engine == "tc" else torch.float32)
    n.call("fdm_cast", n.CastArgs(x=ptr(dy), out=ptr(z), N=Nf, H=H // 2, W=W // 2, C=Cc, upsample=2, op_dtype=dt(z)), stream())
    zr = torch.zeros(Nf, H, W, Cc, device="cuda")
    zr[:, ::2, ::2] = dy
    assert torch.equal(z.float(), zr.to(z.dtype).float())
    if engine == "tc":
        got = conv_native(z, pack(w, n.PACK_TC_DGRAD), Cc, 3, engine=n.CONV_TC).float()
        ref = torch.nn.grad.conv2d_input((Nf, Cc, H, W), w.to(torch.bfloat16).float(),
                                         dy.to(torch.bfloat16).float().permute(0, 3, 1, 2), stride=2, padding=1)
        tol = 6e-3
    else:
        got = conv_native(z, pack(w, n.PACK_SIMT_DGRAD), Cc, 3, engine=n.CONV_SIMT, want="f32")
        ref = torch.nn.grad.conv2d_input((Nf, Cc, H, W), w, dy.permute(0, 3, 1, 2), stride=2, padding=1)
        tol = 1e-5
    assert rel(got, ref.permute(0, 2, 3, 1)) <= tol


def run_wgrad(a, dy, Cw, k, stride, engine):
    n = N_()
    Nf, H, W, Cc = a.shape
    Co = dy.shape[-1]
    dw = torch.full((Co, Cw, k, k), float("nan"), device="cuda")
    db, db2 = torch.empty(Co, device="cuda"), torch.empty(Co, device="cuda")
    args = n.ConvWgradArgs(a=ptr(a), dy=ptr(dy), dw=ptr(dw), dbias=ptr(db), dbias2=ptr(db2), workspace=None, workspace_bytes=0,
                           N=Nf, Hin=H, Win=W, C=Cc, Cw=Cw, Cout=Co, ksize=k, stride=stride, a_dtype=dt(a), dy_dtype=dt(dy),
                           engine=engine)
    nbytes = n.lib().fdm_conv_wgrad_workspace(C.byref(args))
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    args.workspace, args.workspace_bytes = ptr(ws), nbytes
    n.call("fdm_conv_wgrad", args, stream())
    torch.cuda.synchronize()
    return dw, db, db2


WGRAD_CASES = [
    # N, H, W, C (stored), Cw (true), Cout, k, stride
    (5, 32, 32, 64, 64, 64, 3, 1),
    (3, 16, 16, 128, 128, 64, 3, 1),
    (5, 32, 32, 8, 5, 64, 3, 1),     # stem: operand padded 5 -> 8 channels
    (5, 32, 32, 64, 64, 4, 3, 1),    # head
    (4, 32, 32, 64, 64, 64, 3, 2),   # Downsample
    (7, 8, 8, 192, 192, 64, 1, 1),   # 1x1 skip / linear
    (2, 4, 4, 256, 256, 768, 1, 1),  # qkv on a 4x4 map
    (1, 64, 64, 32, 32, 32, 3, 1),
    (2, 128, 128, 128, 128, 128, 3, 1),  # 128-wide rows: one X block per stage, filter rows paired
    (3, 64, 64, 128, 128, 256, 3, 1),
    (3, 32, 32, 256, 256, 128, 1, 1),
    (2, 16, 16, 192, 192, 384, 3, 1),    # ragged 128-channel chunking of both operands
    (40, 32, 32, 64, 64, 64, 3, 1),      # many pixel chunks per split
    (3, 16, 16, 384, 384, 1152, 1, 1),   # qkv linear of the nc = 128 model: 3C = 1152 output columns (two bias-sum column chunks)
]


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_conv_wgrad(case, dtype, engine):
    n = N_()
    if engine == "tc" and dtype == "fp32":
        pytest.skip("tcgen05 wgrad takes bf16 operands")
    Nf, H, W, Cc, Cw, Co, k, stride = case
    torch.manual_seed(3)
    a = torch.randn(Nf, H, W, Cc, device="cuda")
    a[..., Cw:] = 0
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    dy = torch.randn(Nf, Ho, Wo, Co, device="cuda")
    if dtype == "bf16":
        a, dy = a.to(torch.bfloat16), dy.to(torch.bfloat16)
    dw, db, db2 = run_wgrad(a, dy, Cw, k, stride, n.CONV_TC if engine == "tc" else n.CONV_SIMT)
    ref = torch.nn.grad.conv2d_weight(a.float()[..., :Cw].permute(0, 3, 1, 2).double(), (Co, Cw, k, k),
                                      dy.float().permute(0, 3, 1, 2).double(), stride=stride, padding=pad)
    assert rel(dw, ref) <= 2e-5, rel(dw, ref)
    refb = dy.double().sum(dim=(0, 1, 2))
    assert rel(db, refb) <= 1e-5 and torch.equal(db, db2)


def gn_reference(xa, xb, gamma, beta, film, silu, T):
    x = xa if xb is None else torch.cat([xa, xb], dim=-1)
    Nf, HW, Cc = x.shape
    y = F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, 1e-5).permute(0, 2, 1)
    if film is not None:
        sc, sh = film[:, :Cc], film[:, Cc:]
        y = y * (1 + sc.repeat_interleave(T, 0)[:, None]) + sh.repeat_interleave(T, 0)[:, None]
    return F.silu(y) if silu else y


@pytest.mark.parametrize("case", [
    # N, T, HW, Ca, Cb, film, silu, dy_op, dy_f32, draw
    (10, 5, 1024, 64, 0, True, True, True, False, False),
    (10, 5, 256, 128, 64, False, True, True, False, True),
    (6, 3, 64, 192, 0, False, False, True, True, False),
    (4, 2, 16, 256, 256, True, True, True, False, False),
    (5, 5, 1024, 32, 0, False, True, False, True, False),
])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_gn_bwd(case, dtype):
    n = N_()
    Nf, T, HW, Ca, Cb, use_film, silu, use_op, use_f32, use_raw = case
    Cc, B = Ca + Cb, Nf // T
    od = torch.bfloat16 if dtype == "bf16" else torch.float32
    torch.manual_seed(4)
    xa = (torch.randn(Nf, HW, Ca, device="cuda") * 1.5 + 0.3).requires_grad_()
    xb = (torch.randn(Nf, HW, Cb, device="cuda") * 0.7 - 0.2).requires_grad_() if Cb else None
    gamma = (1 + 0.3 * torch.randn(Cc, device="cuda")).requires_grad_()
    beta = (0.2 * torch.randn(Cc, device="cuda")).requires_grad_()
    fs, fo = 2 * Cc + 24, 16
    film_full = (0.3 * torch.randn(B, fs, device="cuda"))
    film = film_full[:, fo:fo + 2 * Cc].clone().requires_grad_() if use_film else None
    dy_op = torch.randn(Nf, HW, Cc, device="cuda").to(od) if use_op else None
    dy_f32 = torch.randn(Nf, HW, Cc, device="cuda") if use_f32 else None
    draw = torch.randn(Nf, HW, Cc, device="cuda").to(od) if use_raw else None
    y = gn_reference(xa, xb, gamma, beta, film, silu, T)
    up = (dy_op.float() if use_op else 0) + (dy_f32 if use_f32 else 0)
    loss = (y * up).sum()
    if use_raw:
        loss = loss + ((xa if xb is None else torch.cat([xa, xb], -1)) * draw.float()).sum()
    loss.backward()

    def stats(x):
        return torch.stack([x.detach().double().sum(1), (x.detach().double() ** 2).sum(1)], dim=-1).contiguous()
    sa, sb = stats(xa), stats(xb) if Cb else None
    pre_a, pre_b = torch.randn_like(xa), torch.randn_like(xb) if Cb else None
    gxa = pre_a.clone()                     # acc_a = 1: accumulates onto existing content
    gxb = torch.full_like(xb, float("nan")) if Cb else None  # acc_b = 0: overwrites
    ab = torch.zeros(Nf, Cc, 2, device="cuda", dtype=torch.float64)
    dgamma, dbeta = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    dfilm = torch.zeros(B, fs, device="cuda") if use_film else None
    a = n.GnBwdArgs(xa=ptr(xa), xb=ptr(xb), stats_a=ptr(sa), stats_b=ptr(sb), gamma=ptr(gamma), beta=ptr(beta),
                    film=ptr(film_full) if use_film else None, dy_op=ptr(dy_op), dy_f32=ptr(dy_f32), draw_op=ptr(draw),
                    gxa=ptr(gxa), gxb=ptr(gxb), ab=ptr(ab), dgamma=ptr(dgamma), dbeta=ptr(dbeta), dfilm=ptr(dfilm),
                    N=Nf, HW=HW, Ca=Ca, Cb=Cb, T=T, film_stride=fs, film_off=fo, silu=int(silu), op_dtype=dt(torch.empty(0, dtype=od)),
                    acc_a=1, acc_b=0, eps=1e-5)
    # fused outputs: operand-dtype copy of the FINAL gradient of xa and its column sums (bias gradient of xa's producer)
    gop = torch.full((Nf, HW, Ca), float("nan"), device="cuda", dtype=od)
    cs, cs2 = torch.zeros(Ca, device="cuda"), torch.ones(Ca, device="cuda")
    a.gop_a, a.cs_a, a.cs2_a = ptr(gop), ptr(cs), ptr(cs2)
    dpass = torch.randn(Nf, HW, Ca, device="cuda")  # fp32 pass-through gradient of xa (identity residual)
    a.dpass_a = ptr(dpass)
    pre_a = pre_a + dpass
    n.call("fdm_gn_bwd", a, stream())
    torch.cuda.synchronize()
    tol = 2e-5
    assert torch.equal(gop.float(), gxa.to(od).float())
    ref_cs = gxa.double().sum(dim=(0, 1))
    assert rel(cs, ref_cs) <= 1e-5 and rel(cs2 - 1, ref_cs) <= 1e-5
    assert rel(gxa - pre_a, xa.grad) <= tol, rel(gxa - pre_a, xa.grad)
    if Cb:
        assert rel(gxb, xb.grad) <= tol
    assert rel(dgamma, gamma.grad) <= tol and rel(dbeta, beta.grad) <= tol
    if use_film:
        assert rel(dfilm[:, fo:fo + 2 * Cc], film.grad) <= tol


@pytest.mark.parametrize("case", [(2, 5, 256, 64), (1, 20, 64, 128), (1, 40, 16, 512), (3, 7, 100, 96)])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_temporal_gn_bwd(case, dtype):
    n = N_()
    B, T, HW, Cc = case
    od = torch.bfloat16 if dtype == "bf16" else torch.float32
    torch.manual_seed(5)
    x = torch.randn(B * T, HW, Cc, device="cuda").requires_grad_()
    gamma = (1 + 0.3 * torch.randn(Cc, device="cuda")).requires_grad_()
    beta = (0.2 * torch.randn(Cc, device="cuda")).requires_grad_()
    dy_op = torch.randn(B * T, HW, Cc, device="cuda").to(od)
    dy_f32 = torch.randn(B * T, HW, Cc, device="cuda")
    xr = x.view(B, T, HW, Cc).permute(0, 2, 3, 1).reshape(B * HW, Cc, T)  # rpe.py:135: [B*D, C, T]
    y = F.group_norm(xr, 32, gamma, beta, 1e-5).view(B, HW, Cc, T).permute(0, 3, 1, 2).reshape(B * T, HW, Cc)
    (y * (dy_op.float() + dy_f32)).sum().backward()
    gx = torch.full_like(x, float("nan"))
    dgamma, dbeta = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    a = n.TemporalGnBwdArgs(x=ptr(x), gamma=ptr(gamma), dy_op=ptr(dy_op), dy_f32=ptr(dy_f32), gx=ptr(gx), dgamma=ptr(dgamma),
                            dbeta=ptr(dbeta), B=B, T=T, HW=HW, C=Cc, op_dtype=dt(dy_op), accumulate=0, eps=1e-5)
    n.call("fdm_temporal_gn_bwd", a, stream())
    torch.cuda.synchronize()
    assert rel(gx, x.grad) <= 2e-5, rel(gx, x.grad)
    assert rel(dgamma, gamma.grad) <= 2e-5 and rel(dbeta, beta.grad) <= 2e-5


@pytest.mark.parametrize("case", [(5, 256, 64, 4), (3, 64, 192, 4), (4, 16, 256, 4), (2, 256, 512, 4), (3, 256, 384, 4), (6, 64, 128, 2)])
def test_attn_spatial_bwd_tcgen05(case):
    """tcgen05 backward kernels fed by the log-sum-exp the tcgen05 FORWARD kernel saves (the training pairing)."""
    n = N_()
    Nf, L, Cc, heads = case
    Fd = Cc // heads
    torch.manual_seed(16)
    qkv = torch.randn(Nf, L, 3 * Cc, device="cuda").to(torch.bfloat16)
    dout = torch.randn(Nf, L, Cc, device="cuda").to(torch.bfloat16)
    out = torch.empty(Nf, L, Cc, device="cuda", dtype=torch.bfloat16)
    lse = torch.full((Nf * heads * L,), float("nan"), device="cuda")
    n.call("fdm_attn_spatial", n.AttnSpatialArgs(qkv=ptr(qkv), out=ptr(out), N=Nf, L=L, C=Cc, heads=heads, qkv_dtype=n.BF16,
                                                 out_dtype=n.BF16, engine=0, lse=ptr(lse)), stream())
    ref_in = qkv.float().requires_grad_()
    q, k, v = ref_in.view(Nf, L, 3, heads, Fd).permute(2, 0, 3, 1, 4)
    w = (q * Fd ** -0.5) @ k.transpose(-1, -2)
    o_ref = (torch.softmax(w, dim=-1) @ v).permute(0, 2, 1, 3).reshape(Nf, L, Cc)
    (o_ref * dout.float()).sum().backward()
    torch.cuda.synchronize()
    assert rel(lse.view(Nf, heads, L), torch.logsumexp(w.detach(), dim=-1)) <= 1e-5
    dqkv = torch.full_like(qkv, float("nan"))
    dsum = torch.empty(Nf * heads * L, device="cuda")
    a = n.AttnSpatialBwdArgs(qkv=ptr(qkv), out=ptr(out), dout=ptr(dout), dqkv=ptr(dqkv), lse=ptr(lse), dsum=ptr(dsum),
                             N=Nf, L=L, C=Cc, heads=heads, dtype=n.BF16, lse_from_forward=1)
    n.call("fdm_attn_spatial_bwd", a, stream())
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, Cc)), ("dk", slice(Cc, 2 * Cc)), ("dv", slice(2 * Cc, 3 * Cc))):
        e = rel(dqkv.float()[..., sl], ref_in.grad[..., sl])
        assert e <= 1.5e-2, (name, e)  # bf16 P / dS operands + bf16 outputs


@pytest.mark.parametrize("case", [(5, 256, 64, 4), (3, 64, 96, 4), (4, 16, 128, 4), (2, 256, 512, 4), (2, 100, 64, 2)])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_attn_spatial_bwd(case, dtype):
    n = N_()
    Nf, L, Cc, heads = case
    Fd = Cc // heads
    od = torch.bfloat16 if dtype == "bf16" else torch.float32
    torch.manual_seed(6)
    qkv = torch.randn(Nf, L, 3 * Cc, device="cuda").to(od)
    dout = torch.randn(Nf, L, Cc, device="cuda").to(od)
    ref_in = qkv.float().requires_grad_()
    q, k, v = ref_in.view(Nf, L, 3, heads, Fd).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q * Fd ** -0.5) @ k.transpose(-1, -2), dim=-1)
    out = (p @ v).permute(0, 2, 1, 3).reshape(Nf, L, Cc)
    (out * dout.float()).sum().backward()
    out_op = out.detach().to(od)
    dqkv = torch.full_like(qkv, float("nan"))
    lse, dsum = torch.empty(Nf * heads * L, device="cuda"), torch.empty(Nf * heads * L, device="cuda")
    a = n.AttnSpatialBwdArgs(qkv=ptr(qkv), out=ptr(out_op), dout=ptr(dout), dqkv=ptr(dqkv), lse=ptr(lse), dsum=ptr(dsum),
                             N=Nf, L=L, C=Cc, heads=heads, dtype=dt(qkv))
    n.call("fdm_attn_spatial_bwd", a, stream())
    torch.cuda.synchronize()
    e = rel(dqkv.float(), ref_in.grad)
    assert e <= (1e-2 if dtype == "bf16" else 2e-5), e


@pytest.mark.parametrize("case", [(2, 5, 256, 64, 4, True), (1, 20, 64, 128, 4, True), (1, 40, 40, 96, 4, False),
                                  (2, 11, 16, 512, 4, True), (1, 3, 1024, 64, 4, False)])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_attn_temporal_bwd(case, dtype):
    n = N_()
    B, T, HW, Cc, heads, use_mask = case
    Fd = Cc // heads
    od = torch.bfloat16 if dtype == "bf16" else torch.float32
    torch.manual_seed(7)
    qkv = torch.randn(B * T, HW, 3 * Cc, device="cuda").to(od)
    dout = torch.randn(B * T, HW, Cc, device="cuda").to(od)
    R = [(0.5 * torch.randn(B, T, T, Cc, device="cuda")).requires_grad_() for _ in range(3)]
    Rq, Rk, Rv = R
    mask = (torch.rand(B, T, device="cuda") > 0.4).float() if use_mask else None
    ref_in = qkv.float().requires_grad_()
    x = ref_in.view(B, T, HW, 3, heads, Fd).permute(3, 0, 2, 4, 1, 5)  # 3 B D H T F
    q, k, v = x[0] * Fd ** -0.5, x[1], x[2]
    w = q @ k.transpose(-1, -2)
    w = w + torch.einsum("bdhtf,btshf->bdhts", q, Rk.view(B, T, T, heads, Fd))
    w = w + torch.einsum("bdhtf,btshf->bdhts", k * Fd ** -0.5, Rq.view(B, T, T, heads, Fd)).transpose(-1, -2)
    if use_mask:
        m = mask.view(B, 1, T)
        allowed = m * m.transpose(1, 2) + (1 - m) * (1 - m.transpose(1, 2))
        w = w.masked_fill((allowed == 0).view(B, 1, 1, T, T), float("-inf"))
    p = torch.softmax(w, dim=-1)
    o = p @ v + torch.einsum("bdhts,btshf->bdhtf", p, Rv.view(B, T, T, heads, Fd))
    out = o.permute(0, 3, 1, 2, 4).reshape(B * T, HW, Cc)  # B T D H F
    (out * dout.float()).sum().backward()
    out_op = out.detach().to(od)
    dqkv = torch.full_like(qkv, float("nan"))
    dR = [torch.zeros(B, T, T, Cc, device="cuda") for _ in range(3)]
    rows = B * heads * T * HW
    lse, dsum = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    a = n.AttnTemporalBwdArgs(qkv=ptr(qkv), out=ptr(out_op), Rq=ptr(Rq), Rk=ptr(Rk), Rv=ptr(Rv), mask=ptr(mask), dout=ptr(dout),
                              dqkv=ptr(dqkv), dRq=ptr(dR[0]), dRk=ptr(dR[1]), dRv=ptr(dR[2]), lse=ptr(lse), dsum=ptr(dsum),
                              B=B, T=T, HW=HW, C=Cc, heads=heads, dtype=dt(qkv))
    n.call("fdm_attn_temporal_bwd", a, stream())
    torch.cuda.synchronize()
    tol = 1e-2 if dtype == "bf16" else 3e-5
    e = rel(dqkv.float(), ref_in.grad)
    assert e <= tol, ("dqkv", e)
    for name, got, ref in zip(("dRq", "dRk", "dRv"), dR, R):
        e = rel(got, ref.grad)
        assert e <= tol, (name, e)


def test_rpe_hidden_bwd_and_grouped_linear_bwd():
    n = N_()
    torch.manual_seed(8)
    B, T, ts = 3, 7, 200
    fi = torch.stack([torch.sort(torch.randperm(50)[:T]).values for _ in range(B)]).cuda()
    te = torch.randn(B, ts, device="cuda").requires_grad_()
    nets = []
    for Cc, off in [(64, 8), (96, 100)]:
        nets.append(dict(C=Cc, off=off, wd=torch.randn(Cc, 3, device="cuda").requires_grad_(),
                         bd=torch.randn(Cc, device="cuda").requires_grad_(), dh=torch.randn(B, T, T, Cc, device="cuda")))
    d = (fi[:, :, None] - fi[:, None, :]).float()
    feats = torch.stack([torch.log(1 + d.clamp(min=0)), torch.log(1 + (-d).clamp(min=0)), (d == 0).float()], dim=-1)
    loss = 0
    for nt in nets:
        e = te[:, nt["off"]:nt["off"] + nt["C"]].view(B, 1, 1, -1) + feats @ nt["wd"].t() + nt["bd"]
        loss = loss + (F.silu(e) * nt["dh"]).sum()
    loss.backward()
    dte = torch.zeros(B, ts, device="cuda")
    probs = []
    for nt in nets:
        nt["dwd"], nt["dbd"] = torch.zeros(nt["C"], 3, device="cuda"), torch.zeros(nt["C"], device="cuda")
        probs.append(n.RpeHiddenBwdProblem(wd=ptr(nt["wd"]), bd=ptr(nt["bd"]), dhidden=ptr(nt["dh"]), dwd=ptr(nt["dwd"]),
                                           dbd=ptr(nt["dbd"]), C=nt["C"], te_off=nt["off"]))
    pa = dev_array(probs)
    n.call("fdm_rpe_hidden_bwd", n.RpeHiddenBwdArgs(te=ptr(te), frame_indices=ptr(fi), problems=ptr(pa), dte=ptr(dte), B=B, T=T,
                                                    te_stride=ts, count=2, max_C=96, dhidden_dtype=n.F32), stream())
    torch.cuda.synchronize()
    assert rel(dte, te.grad) <= 2e-5
    for nt in nets:
        assert rel(nt["dwd"], nt["wd"].grad) <= 2e-5 and rel(nt["dbd"], nt["bd"].grad) <= 2e-5

    # grouped linear backward: two problems sharing x (one with SiLU on the input), M = 3 and a 13-row problem
    M, K = 3, 256
    x = torch.randn(M, K, device="cuda").requires_grad_()
    specs = [(128, 1), (200, 0)]
    ws = [torch.randn(No, K, device="cuda").requires_grad_() for No, _ in specs]
    bs = [torch.randn(No, device="cuda").requires_grad_() for No, _ in specs]
    ldy = 400
    dy = torch.randn(M, ldy, device="cuda")
    loss, off = 0, 0
    offs = []
    for (No, s), w, b in zip(specs, ws, bs):
        y = F.linear(F.silu(x) if s else x, w, b)
        loss = loss + (y * dy[:, off:off + No]).sum()
        offs.append(off)
        off += No
    loss.backward()
    parts = torch.full((2, M, K), float("nan"), device="cuda")
    dws = [torch.empty_like(w) for w in ws]
    dbs = [torch.empty_like(b) for b in bs]
    probs = [n.LinearBwdProblem(x=ptr(x), w=ptr(w), dy=dy.data_ptr() + 4 * o, dw=ptr(dw), db=ptr(db), dx_part=parts[i].data_ptr(),
                                M=M, K=K, Nout=No, ldx=K, ldy=ldy, silu_in=s)
             for i, ((No, s), w, dw, db, o) in enumerate(zip(specs, ws, dws, dbs, offs))]
    pa = dev_array(probs)
    n.call("fdm_grouped_linear_bwd", n.GroupedLinearBwdArgs(problems=ptr(pa), count=2, max_M=M, max_Nout=200, max_K=K), stream())
    dx = torch.randn(M, K, device="cuda")
    pre = dx.clone()
    n.call("fdm_sum_parts", n.SumPartsArgs(parts=ptr(parts), out=ptr(dx), part_stride=M * K, n=M * K, count=2, accumulate=1), stream())
    torch.cuda.synchronize()
    assert rel(dx - pre, x.grad) <= 2e-5
    for dw, w, db, b in zip(dws, ws, dbs, bs):
        assert rel(dw, w.grad) <= 2e-5 and rel(db, b.grad) <= 2e-5


def test_accum_pool_and_nchw_to_nhwc():
    n = N_()
    torch.manual_seed(9)
    Nf, H, W, Cc = 3, 8, 8, 64
    src = torch.randn(Nf, 2 * H, 2 * W, Cc, device="cuda").to(torch.bfloat16)
    dst = torch.randn(Nf, H, W, Cc, device="cuda")
    pre = dst.clone()
    n.call("fdm_accum", n.AccumArgs(src=ptr(src), dst=ptr(dst), N=Nf, H=H, W=W, C=Cc, pool=1, src_dtype=n.BF16, accumulate=1), stream())
    ref = pre + F.avg_pool2d(src.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * 4
    assert rel(dst, ref) <= 1e-6
    src2 = torch.randn(Nf, H, W, Cc, device="cuda")
    n.call("fdm_accum", n.AccumArgs(src=ptr(src2), dst=ptr(dst), N=Nf, H=H, W=W, C=Cc, pool=0, src_dtype=n.F32, accumulate=0), stream())
    assert torch.equal(dst, src2)
    g = torch.randn(Nf, 4, H, W, device="cuda")
    out = torch.full((Nf, H, W, 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    n.call("fdm_nchw_to_nhwc", n.NchwToNhwcArgs(src=ptr(g), dst=ptr(out), N=Nf, C=4, H=H, W=W, Cpad=8, op_dtype=n.BF16), stream())
    torch.cuda.synchronize()
    assert torch.equal(out[..., :4].float(), g.permute(0, 2, 3, 1).to(torch.bfloat16).float()) and float(out[..., 4:].abs().max()) == 0


@pytest.mark.parametrize("case", [(5, 32, 32, 64), (3, 8, 8, 192), (1, 4, 4, 1024), (7, 1, 1, 96)])
def test_cast_with_fused_bias_gradient(case):
    """fdm_cast with colsum: the operand-dtype copy of a gradient and its column sums (= the conv's bias gradient) in one pass."""
    n = N_()
    Nf, H, W, Cc = case
    torch.manual_seed(10)
    g = torch.randn(Nf, H, W, Cc, device="cuda")
    out = torch.empty(Nf, H, W, Cc, device="cuda", dtype=torch.bfloat16)
    cs, cs2 = torch.zeros(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    n.call("fdm_cast", n.CastArgs(x=ptr(g), out=ptr(out), N=Nf, H=H, W=W, C=Cc, upsample=0, op_dtype=n.BF16, colsum=ptr(cs),
                                  colsum2=ptr(cs2)), stream())
    torch.cuda.synchronize()
    assert torch.equal(out, g.to(torch.bfloat16))
    ref = g.double().sum(dim=(0, 1, 2))
    assert rel(cs, ref) <= 1e-5 and rel(cs2 - 1, ref) <= 1e-5
