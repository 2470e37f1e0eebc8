"""One fused-LSTM-pair launch at M rows (for ncu captures): python scripts/lstm_once.py M"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench_roofline import lstm_pair
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
print(lstm_pair(M, 368, 256, torch.device("cuda", 0), reps=5, x3=1))
