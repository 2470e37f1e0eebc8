"""A few EAGER train iterations (forward, loss, backward, Adam) of a bench workload, for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <count> --csv --log-file out.csv \\
        python scripts/one_iter.py c4 256 [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, model_config
from marlclassification_b200.config import ModelConfig
from marlclassification_b200.core import EpisodeSampler
from marlclassification_b200.training.optim import FlatAdam

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = WORKLOADS[wl]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else w["B"]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model, marl, env = ModelConfig(**model_config(w)).build_marl(w["na"])
model.to(dev)
sampler = EpisodeSampler(marl, env, w["T"])
img = torch.rand(nb, w["C"], w["H"], w["W"], device=dev)
y = torch.randint(w["nc"], (nb,), device=dev)
eng = sampler.engine_for(img)
opt = FlatAdam(model, 1e-4)
for i in range(iters):
    eng.forward(img)
    eng.loss(y)
    eng.backward(img)
    opt.step()
torch.cuda.synchronize()
print("launches per iteration:", eng.launches, "+ 2 (Adam)")
