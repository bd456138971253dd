#!/usr/bin/env python3
"""Sweep the token count of the Gemma gate/up GEMM shape: separates weight streaming from the
activation operand."""
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import capi

lib = capi.load_library()
dev = torch.device("cuda:0")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
N, K = 32768, 2048
Ws = [torch.empty((N, K), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02) for _ in range(4)]
for use2 in (0, 1):
    capi.check(lib.blurr_set_global_option(b"gemm_use_2cta", use2))
    for epi, name in ((capi.EPI_STORE, "store"), (capi.EPI_GEGLU, "geglu")):
        for T in (16, 64, 128, 192, 256, 276):
            X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
            out = torch.empty((T, N), device=dev, dtype=torch.bfloat16)
            ldo = N // 2 if epi == capi.EPI_GEGLU else N

            def launch(i):
                return lib.blurr_op_gemm_async(sp, C.c_void_p(Ws[i % 4].data_ptr()), N, K, 0, C.c_void_p(X.data_ptr()), T, K,
                                               epi, 1, None, C.c_void_p(out.data_ptr()), ldo, None)
            for i in range(4):
                capi.check(launch(i))
            torch.cuda.synchronize()
            pairs = []
            for i in range(20):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); capi.check(launch(i)); e.record()
                pairs.append((s, e))
            torch.cuda.synchronize()
            ms = statistics.fmean(s.elapsed_time(e) for s, e in pairs)
            print(f"2cta={use2} {name:5s} T={T:4d}: {ms * 1e3:7.1f} us  {N * K * 2 / ms / 1e6:6.0f} GB/s", flush=True)
