"""Debug: where a conv_halo launch spends its cycles (per-CTA counters).  python tools/trace_halo.py N H W C0 Cout [resid]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"), os.path.join(ROOT, "tests")]
import torch
from improved_diffusion import _native as N_
from test_gpu_kernels import run_conv

N, H, W, C0, Co = map(int, sys.argv[1:6])
use_resid = "resid" in sys.argv[6:]
kw = dict(want_stats="nostats" not in sys.argv[6:], want_op="op" in sys.argv[6:], want_f32="nof32" not in sys.argv[6:])
x = torch.randn(N, H, W, C0, device="cuda").to(torch.bfloat16)
w = torch.randn(Co, C0, 3, 3, device="cuda") / (C0 * 9) ** 0.5
b = torch.randn(Co, device="cuda")
resid = torch.randn(N, H, W, Co, device="cuda") if use_resid else None
lib = N_.lib()
for rep in range(3):
    tr = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
    lib.fdm_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    run_conv(x, w, b, engine=N_.CONV_TC, resid=resid, **kw)
lib.fdm_debug_set_trace(ctypes.c_void_p(0))
t = tr.cpu().float()
t = t[t[:, 6] > 0]
names = ["MMA warp total", "  waiting for a free accumulator", "  waiting for operands (TMA)", "epilogue total", "  waiting for the accumulator", "  statistics phase", "items"]
for i, n in enumerate(names):
    print(f"{n:36s} mean {t[:, i].mean().item():10.0f}   max {t[:, i].max().item():10.0f}   (clk)")
items = t[:, 6].mean().item()
print(f"per item: MMA issue+wait {t[:,0].mean().item()/items:.0f} clk, epilogue busy {(t[:,3]-t[:,4]).mean().item()/items:.0f} clk")
