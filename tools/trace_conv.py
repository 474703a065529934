"""Debug: per-CTA phase timeline of one conv_tc launch (globaltimer stamps).  python tools/trace_conv.py N H W C0 Cout k [resid]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"), os.path.join(ROOT, "tests")]
import torch
from improved_diffusion import _native as N_
from test_gpu_kernels import run_conv

N, H, W, C0, Co, k = map(int, sys.argv[1:7])
use_resid = len(sys.argv) > 7
x = torch.randn(N, H, W, C0, device="cuda").to(torch.bfloat16)
w = torch.randn(Co, C0, k, k, device="cuda") / (C0 * k * k) ** 0.5
b = torch.randn(Co, device="cuda")
resid = torch.randn(N, H, W, Co, device="cuda") if use_resid else None
lib = N_.lib()
ncta = (N * H * W + 127) // 128 * max(1, (Co + 127) // 128 if Co % 128 == 0 else (Co + 63) // 64)
for rep in range(3):
    tr = torch.zeros(ncta * 4, 8, dtype=torch.int64, device="cuda")
    lib.fdm_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    run_conv(x, w, b, engine=N_.CONV_TC, resid=resid)
    e1.record()
    torch.cuda.synchronize()
lib.fdm_debug_set_trace(ctypes.c_void_p(0))
t = tr.cpu()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["prologue", "first operands", "main loop (to last MMA issue)", "accumulator ready", "TMEM->smem staging", "store/stats", ]
print(f"CTAs {len(t)}  kernel span {(t[:, 6].max() - t0).item() / 1e3:.1f} us  (event incl. launch+sync: {e0.elapsed_time(e1) * 1e3:.1f} us)")
d = (t[:, 1:7] - t[:, 0:6]).float()
for i, n in enumerate(names):
    print(f"  {n:32s} mean {d[:, i].mean().item() / 1e3:7.2f} us   p90 {d[:, i].quantile(0.9).item() / 1e3:7.2f} us")
life = (t[:, 6] - t[:, 0]).float()
print(f"  CTA lifetime mean {life.mean().item() / 1e3:.2f} us; start spread: p50 {((t[:,0]-t0).float().median().item())/1e3:.1f} us")
