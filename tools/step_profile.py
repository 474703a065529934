"""Per-launch HOT timing of one sampler step: every forward launch of the plan repeated back to back (operands L2-warm), CUDA
events on the launching stream.  python tools/step_profile.py [workload] > profiles/r02_step_profile_<workload>.txt"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200")):
    sys.path.insert(0, p)
import torch as th  # noqa: E402

import bench  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4-sampling"]
dev = th.device("cuda", 0)
model, diffusion, sd = bench.build_native(wl["over"], dev)
over, B, K = wl["over"], wl["B"], wl["K"]
S = over["image_size"]
eng = model.engine()
P = eng.plan_for(B, K, S, S, dev)
batch = {k: v.to(dev) for k, v in bench.synthetic_batch(over, B, K, wl["n_obs"], wl["video_len"], seed=0).items()}
eng.load_conditioning(P, batch["x0"], batch["frame_indices"], batch["obs_mask"], batch["latent_mask"])
P.set_t_source(None)
P.x_view.normal_()
P.t_view.fill_(500.0)
st = th.cuda.current_stream(dev)
P.run(st.cuda_stream)
st.synchronize()
sp = C.c_void_p(st.cuda_stream)
rows, total = [], 0.0
for i, ((name, fn, ref), s_) in enumerate(zip(P.calls, P._structs)):
    reps = 20
    fn(ref, sp)
    st.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(ref, sp)
    e1.record()
    st.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    total += us
    d = ""
    if name == "fdm_conv":
        d = f"N={s_.N} {s_.Hin}x{s_.Win} C0={s_.C0} C1={s_.C1} Cout={s_.Cout} k={s_.ksize} s={s_.stride} eng={'tc' if s_.engine == 0 else 'simt'} f32={int(bool(s_.y_f32))} op={int(bool(s_.y_op))} res={int(bool(s_.resid))} st={int(bool(s_.stats))}"
    elif name == "fdm_gn_apply":
        d = f"N={s_.N} HW={s_.HW} Ca={s_.Ca} Cb={s_.Cb} film={int(bool(s_.film))} f32={int(bool(s_.out_f32))} raw={int(bool(s_.raw_op))}"
    elif name in ("fdm_attn_temporal", "fdm_temporal_gn"):
        d = f"B={s_.B} T={s_.T} HW={s_.HW} C={s_.C}"
    elif name == "fdm_attn_spatial":
        d = f"N={s_.N} L={s_.L} C={s_.C}"
    side = "side" if P.side_begin <= i < P.side_end else "    "
    rows.append((i, name, us, side, d))
print(f"# {sys.argv[1] if len(sys.argv) > 1 else 'cfg4-sampling'}: {len(rows)} launches, sum of hot per-launch times {total:.0f} us")
for i, name, us, side, d in rows:
    print(f"{i:4d} {side} {name:24s} {us:8.1f} us  {d}")
