"""Phase timeline of fdm_norm_linear (clock64 stamps written through the debug hook fdm_debug_nl_trace).
python tools/nl_trace.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from improved_diffusion import _native as N_  # noqa: E402
import test_gpu_kernels as K  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
L = N_.lib()
L.fdm_debug_nl_trace.argtypes = [C.c_void_p]
L.fdm_debug_nl_trace.restype = None
GHZ = 1.965
NAMES = {0: "raw_full", 1: "A produced", 2: "mma: a_full seen", 3: "mma: issued", 4: "epi: t_full", 5: "epi: tile done"}


def run(label, B, T, HW, a_mode):
    C_, eps, Cout = 128, 1e-5, 384
    M = B * T * HW
    x, gamma, beta, w, bias, g = K._nl_inputs(B, T, HW, Cout, 1)
    wq = w.to(torch.bfloat16).contiguous()
    stats = K._frame_stats(x)
    tstats = torch.zeros(B, HW, 32, 2, device="cuda")
    yop = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    a = N_.NormLinearArgs(x=x.data_ptr(), stats=stats.data_ptr(), tstats=tstats.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                          w=wq.data_ptr(), bias=bias.data_ptr(), y_op=yop.data_ptr(), B=B, T=T, HW=HW, K=C_, Cout=Cout, a_mode=a_mode,
                          eps=eps)
    for _ in range(3):
        N_.call("fdm_norm_linear", a, st)
    torch.cuda.synchronize()
    tr = torch.zeros(148, 64, dtype=torch.int64, device="cuda")
    L.fdm_debug_nl_trace(C.c_void_p(tr.data_ptr()))
    N_.call("fdm_norm_linear", a, st)
    torch.cuda.synchronize()
    L.fdm_debug_nl_trace(None)
    t = tr.cpu()
    dur = (t[:, 3] - t[:, 0]).float() / GHZ / 1e3
    used = t[:, 3] > 0
    slow = int(torch.argmax(dur * used))
    print(f"== {label}: CTA life (after setup) min {dur[used].min():.2f} / mean {dur[used].mean():.2f} / max {dur[used].max():.2f} us; slowest CTA {slow}")
    for cta in (0, slow):
        row = t[cta]
        t0 = int(row[0])
        us = lambda v: (int(v) - t0) / GHZ / 1e3
        out = [f"cta {cta}: pdl_wait done {us(row[1]):.2f}  W landed {us(row[2]):.2f}  end {us(row[3]):.2f}"]
        for it in range(4):
            seg = row[8 + 8 * it: 8 + 8 * it + 6]
            if int(seg[3]) == 0:
                break
            out.append(f"   tile {it}: " + "  ".join(f"{NAMES[k]} {us(seg[k]):.2f}" for k in range(6) if int(seg[k]) > 0))
        print("\n".join(out))


run("qkv spatial (a1) 16x16", 8, 20, 256, 1)
run("qkv temporal (a2) 16x16", 8, 20, 256, 2)
run("qkv spatial (a1) 8x8", 8, 20, 64, 1)
run("qkv temporal (a2) 4x4", 8, 20, 16, 2)
