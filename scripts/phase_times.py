"""Device time of each phase of the train step (forward / loss / backward / adam), each captured
as its own CUDA graph and replayed (so host launch overhead is excluded)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, model_config
from marlclassification_b200.config import ModelConfig
from marlclassification_b200.core import EpisodeSampler
from marlclassification_b200.training.optim import FlatAdam

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = WORKLOADS[wl]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else w["B"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model, marl, env = ModelConfig(**model_config(w)).build_marl(w["na"])
model.to(dev)
for a in sys.argv[3:]:
    if a == "fp32": model.use_tc = False
    if a == "nochains": model.use_chains = False
    if a == "tf32": model.precision = "tf32"
sampler = EpisodeSampler(marl, env, w["T"])
img = torch.rand(nb, w["C"], w["H"], w["W"], device=dev)
y = torch.randint(w["nc"], (nb,), device=dev)
eng = sampler.engine_for(img)
opt = FlatAdam(model, 1e-4)
def bwd(stop):
    def f():
        eng._L.marlc_engine_debug_stop(eng._h, stop)
        eng.backward(img)
        eng._L.marlc_engine_debug_stop(eng._h, 0)
    return f
def fwd(stop):
    def f():
        eng._L.marlc_engine_debug_stop(eng._h, stop)
        eng.forward(img)
        eng._L.marlc_engine_debug_stop(eng._h, 0)
    return f
phases = {"fwd:pre": fwd(11), "fwd:pre+lstm": fwd(12), "fwd:pre+lstm+g0": fwd(13), "fwd:loop": fwd(14),
          "forward": lambda: eng.forward(img), "loss": lambda: eng.loss(y), "bwd:heads": bwd(1), "bwd:heads+pre": bwd(21), "bwd:heads+pre+dx": bwd(22), "bwd:heads+sweep": bwd(2),
          "backward": lambda: eng.backward(img), "adam": lambda: opt.step()}
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        for f in phases.values():
            f()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
tot = 0.0
for name, f in phases.items():
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        f()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 20
    tot += 0.0 if ":" in name else t
    print(f"{name:18s} {t * 1e3:9.1f} us   launches {eng.launches.get(name, 2)}")
print(f"total     {tot * 1e3:9.1f} us  -> {nb / tot * 1e3:.0f} image-episodes/s  ({wl}, batch {nb})")
