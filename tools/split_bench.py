"""Prototype: one diffusion step of the cfg4 shard as ONE plan (B videos) vs TWO/FOUR concurrent sub-batch plans on separate streams
inside one CUDA graph (the videos of a batch are independent).  python tools/split_bench.py [parts]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch as th
import bench
from improved_diffusion import _native as N_
over = dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1, diffusion_steps=1000)
B, K = 8, 20
dev = th.device("cuda:0")
model, diffusion, _ = bench.build_native(over, dev)
model.eval()
eng = model.engine()
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(over, B, K, 10, 300, seed=1).items()}
tb = diffusion._tables(dev)
flush = th.empty(192 * 1024 * 1024, dtype=th.uint8, device=dev)

def build(parts):
    n = B // parts
    plans, steps, nbufs = [], [], []
    for i in range(parts):
        P = eng.plan_for(n, K, 32, 32, dev, slot=i)
        sl = slice(i * n, (i + 1) * n)
        eng.load_conditioning(P, batch["x0"][sl], batch["frame_indices"][sl], batch["obs_mask"][sl], batch["latent_mask"][sl])
        P.set_t_source(tb["model_t"])
        nb = th.randn(n, K, 4, 32, 32, device=dev)
        st = N_.DdpmStepArgs(x=P.ptr(P.x), eps=P.ptr(P.eps), noise=nb.data_ptr(), coef=tb["step"].data_ptr(), t=P.ptr(P.t_index),
                             sample=P.ptr(P.x), pred_xstart=None, per_video=K * 4 * 32 * 32, B=n, clip=1)
        P.x_view.normal_(); P.t_index_view.fill_(500)
        plans.append(P); steps.append(st); nbufs.append(nb)
    streams = [None] + [th.cuda.Stream(dev) for _ in range(parts - 1)]
    ev_f, ev_j = th.cuda.Event(), [th.cuda.Event() for _ in range(parts - 1)]

    def body():
        main = th.cuda.current_stream(dev)
        ev_f.record(main)
        for i, (P, st) in enumerate(zip(plans, steps)):
            s = main if i == 0 else streams[i]
            if i > 0:
                s.wait_event(ev_f)
            with th.cuda.stream(s):
                P.run(s.cuda_stream)
                N_.call("fdm_ddpm_step", st, s.cuda_stream)
            if i > 0:
                ev_j[i - 1].record(s)
        for e in ev_j:
            main.wait_event(e)
    body(); th.cuda.synchronize()
    g = th.cuda.CUDAGraph()
    with th.cuda.graph(g):
        body()
    return g

for parts in [1, 2, 4, 1, 2]:
    g = build(parts)
    for _ in range(5):
        g.replay()
    th.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); th.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"parts={parts}: median {ts[len(ts)//2]:.3f} ms  mean {sum(ts)/len(ts):.3f} ms per diffusion step of {B*K} frames")
