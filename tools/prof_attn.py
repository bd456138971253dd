#!/usr/bin/env python3
"""Run the attention kernels stand-alone a few times (for ncu captures / timing)."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from helpers import op_joint_attention, op_siglip_attention
from blurr_b200 import capi
if os.environ.get("ATTN_TC"):
    capi.check(capi.load_library().blurr_set_global_option(b"attn_tc", int(os.environ["ATTN_TC"])))

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).to(dev)
q = rnd(B * 276, 2048); kc = rnd(B, 281, 256); vc = rnd(B, 281, 256)
mask = torch.zeros(B, 277, 277, dtype=torch.bfloat16, device=dev)
qa = rnd(B * 4, 2048); maska = torch.zeros(B, 4, 281, dtype=torch.bfloat16, device=dev)
qkv = rnd(B * 256, 3456)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return statistics.median(ts)


print("prefill  us:", timeit(lambda: op_joint_attention(False, q, 276, 0, kc, vc, 277, mask, B, 8)))
qp = rnd(B * 1, 2048)
print("fewq q1  us:", timeit(lambda: op_joint_attention(True, qp, 1, 276, kc, vc, 277, mask, B, 8)))
print("fewq q4  us:", timeit(lambda: op_joint_attention(True, qa, 4, 0, kc, vc, 281, maska, B, 8)))
print("siglip   us:", timeit(lambda: op_siglip_attention(qkv, B, 256, 16, 1152)))
