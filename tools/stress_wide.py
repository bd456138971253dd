#!/usr/bin/env python3
"""Stress the two-tile GEMM variant: many launches in the patterns the unit tests use."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from blurr_b200 import capi
from helpers import op_gemm
lib = capi.load_library()
capi.check(lib.blurr_set_global_option(b"gemm_wide", 1))      # the variant is off by default
tot = 0
for (T, N, K) in [(276, 2048, 16384), (276, 2560, 2048), (276, 2048, 2048), (270, 512, 4352), (276, 4096, 2048)]:
    g = torch.Generator().manual_seed(K + N)
    W = (torch.randn((N, K), generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
    X = torch.randn((T, K), generator=g).to(torch.bfloat16).cuda()
    ref = (X.float() @ W.float().t()).to(torch.bfloat16).float()
    nbad = 0
    for it in range(30):
        # a different-shaped GEMM right before, as in the test-suite
        Wd = (torch.randn((1152, 640), generator=g) / 25).to(torch.bfloat16).cuda()
        Xd = torch.randn((256, 640), generator=g).to(torch.bfloat16).cuda()
        op_gemm(Wd, Xd, capi.EPI_STORE)
        b = op_gemm(W, X, capi.EPI_STORE)
        bad = (b.float() - ref).abs() > 0.03
        nbad += int(bad.any())
    tot += nbad
    print(f"T={T} N={N} K={K}: {nbad}/30 wrong", flush=True)
print("TOTAL wrong", tot)
