"""In-kernel timeline of the wide feature-extractor role (build with MARLC_NVCC_EXTRA=-DMARLC_CNN_TRACE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, model_config
from marlclassification_b200.config import ModelConfig
from marlclassification_b200.core import EpisodeSampler
w = WORKLOADS["c4"]; nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
model, marl, env = ModelConfig(**model_config(w)).build_marl(w["na"]); model.to(dev)
sampler = EpisodeSampler(marl, env, 2)
img = torch.rand(nb, 3, 256, 256, device=dev)
eng = sampler.engine_for(img)
eng.forward(img); torch.cuda.synchronize()
