"""Print the compiled kernel schedules of a model configuration WITHOUT a GPU (plans compile over a CPU arena; nothing is launched):
python tools/plan_dump.py [cfg2|cfg3|cfg4|cfg5] [bf16|fp32] [--train] [--ops]
Shows launches per entry point, arena sizes, algorithmic FLOPs, and (with --ops) every launch with its dimensions, the stream it
runs on and where the side stream is joined."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "latent-flexible-video-diffusion-modeling_b200"))
import torch as th
from improved_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults

CFG = {"cfg2": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1), 1, 5),
       "cfg3": (dict(image_size=128, in_channels=3, num_channels=128, num_res_blocks=1), 1, 4),   # short clip: the CPU arena is real memory
       "cfg4": (dict(image_size=32, in_channels=4, num_channels=64, num_res_blocks=1), 8, 20),
       "cfg5": (dict(image_size=64, in_channels=4, num_channels=128, num_res_blocks=1), 1, 40)}
args = [a for a in sys.argv[1:] if not a.startswith("--")]
over, B, T = CFG[args[0] if args else "cfg2"]
prec = args[1] if len(args) > 1 else "bf16"
train, show = "--train" in sys.argv, "--ops" in sys.argv
d = model_and_diffusion_defaults()
d.update(over, diffusion_steps=1000)
d["diffusion_space_kwargs"] = dict(diffusion_space="pixel", pre_encoded=False, pre_encoded_stats_dict=None)
model, _ = create_model_and_diffusion(**d)
model.precision = prec
S = over["image_size"]
P = model.engine()._compile(B, T, S, S, th.device("cpu"), train=train)
DIMS = ("N", "B", "T", "HW", "L", "Hin", "Win", "H", "W", "C", "C0", "C1", "Ca", "Cb", "Cw", "Cout", "ksize", "stride", "heads", "count",
        "phases", "upsample", "engine")
print(f"{args[0] if args else 'cfg2'} {prec} B={B} T={T}: params {sum(p.numel() for p in model.parameters()) / 1e6:.2f} M, "
      f"arena {P.arena_bytes / 1e6:.1f} MB, statistics arena {P.stats_bytes / 1e3:.1f} KB")
for title, ops, flops in (("forward", P.ops, P.flops),) + ((("backward", P.bops, P.bflops),) if train else ()):
    cnt = collections.Counter(fn for fn, _, _ in ops)
    print(f"--- {title}: {len(ops)} C-ABI calls, {flops / 1e9:.2f} GFLOP algorithmic")
    print("    " + ", ".join(f"{k} x{v}" for k, v in cnt.most_common()))
    if show:
        for i, (fn, _, f) in enumerate(ops):
            where = "main"
            if title == "forward" and P.side_begin <= i < P.side_end:
                where = "side"
            if title == "backward" and i in P.bside:
                where = "side"
            note = "   <- main stream joins the side stream here" if ((title == "backward" and i == P.bjoin_before) or
                                                                     (title == "forward" and i == P.join_at and P.side_end > P.side_begin)) else ""
            dims = " ".join(f"{k}={f[k]}" for k in DIMS if k in f and not isinstance(f[k], (th.Tensor, tuple)) and f[k] is not None
                            and not hasattr(f[k], "nbytes"))
            print(f"  {i:4d} [{where}] {fn:24s} {dims}{note}")
if train:
    print(f"--- training: {len(P.pack_problems)} weight-pack problems in one launch, parameter-gradient buffer {P.pgrad.numel() * 4 / 1e6:.1f} MB, "
          f"accumulator arena {P.bzero_bytes / 1e3:.1f} KB, side-stream launches {len(P.bside)}, join before op {P.bjoin_before}")
