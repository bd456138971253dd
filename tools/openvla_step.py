#!/usr/bin/env python3
"""One OpenVLA-7B-shaped generate call for profilers: python tools/openvla_step.py [batch] [n_new] [layers]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import openvla

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
cfg = openvla.LlamaShapedConfig(num_layers=L)
dec = openvla.LlamaDecoder.from_state_dict(cfg, openvla.synthetic_llama_state_dict(cfg, dev, 0), dev, max_batch=B)
dec.set_option("use_cuda_graph", int(os.environ.get("GRAPH", "0")))
x = (torch.randn((B, 281, cfg.hidden), device=dev) * 0.5).to(torch.bfloat16)
for _ in range(int(os.environ.get("REPS", "2"))):
    ids = dec.generate(x, N)
dec.check()
torch.cuda.synchronize()
print("ids", ids[0].tolist(), "launches", dec.last_launch_count)
