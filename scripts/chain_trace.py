"""In-kernel timelines of the fused chain kernels (build with MARLC_NVCC_EXTRA=-DMARLC_CHAIN_TRACE):
one short forward + backward of a workload; prints CTA 0's cycle stamps of the traced kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, model_config
from marlclassification_b200.config import ModelConfig
from marlclassification_b200.core import EpisodeSampler

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model, marl, env = ModelConfig(**model_config(w)).build_marl(w["na"])
model.to(dev)
sampler = EpisodeSampler(marl, env, 3)
img = torch.rand(w["B"], w["C"], w["H"], w["W"], device=dev)
y = torch.randint(w["nc"], (w["B"],), device=dev)
eng = sampler.engine_for(img)
for _ in range(2):
    eng.forward(img)
    eng.loss(y)
    eng.backward(img)
    torch.cuda.synchronize()
