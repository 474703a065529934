"""Training-step timing (native vs autograd path) on one GPU: python tools/train_bench.py [precision]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as th
import bench
dev = th.device("cuda:0")
th.cuda.set_device(dev)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for wl in (sys.argv[2:] or ["cfg2-train", "cfg3-train"]):
    print(json.dumps(bench.bench_train(dev, 1, prec, wl)))
