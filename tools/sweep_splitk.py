#!/usr/bin/env python3
"""Sweep (token chunk width, K slices) of the split-K GEMMs at batch 1: GEMM time (CUDA events, rotating weights) plus
the bytes the consumer has to re-read, to calibrate Run::plan_partial (engine.cu)."""
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
SHAPES = [("vlm o", 2048, 2048, 276), ("vlm qkv", 2560, 2048, 276), ("vlm down", 2048, 16384, 276),
          ("siglip out", 1152, 1152, 256), ("siglip fc2", 1152, 4352, 256)]
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
for name, N, K, T in SHAPES:
    if ONLY and ONLY not in name:
        continue
    nbuf = 4
    Ws = [torch.empty((N, K), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02) for _ in range(nbuf)]
    X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
    part = torch.empty((16 * T * N,), device=dev, dtype=torch.float32)
    tiles = N // 128
    for bn in (0, 144, 128, 96, 64):
        if bn and (T + bn - 1) // bn < 2:
            continue
        chunks = 1 if bn == 0 else (T + bn - 1) // bn
        if bn in (144, 96) and T == 256 or bn in (128, 64) and T == 276:
            continue
        for S in (1, 2, 3, 4, 5, 7, 9, 12, 16):
            if tiles * chunks * S > 148 or S > K // 128:
                continue
            capi.check(lib.blurr_set_global_option(b"op_gemm_bn", bn))

            def launch(i):
                return lib.blurr_op_gemm_async(sp, C.c_void_p(Ws[i % nbuf].data_ptr()), N, K, 0, C.c_void_p(X.data_ptr()), T, K,
                                               capi.EPI_PARTIAL, S, None, None, N, C.c_void_p(part.data_ptr()))
            for i in range(3):
                s_used = launch(i)
                if s_used < 0:
                    break
            if s_used < 0:
                continue
            torch.cuda.synchronize()
            ts = []
            for i in range(15):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); launch(i); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            print(f"{name:11s} T={T} bn={bn:3d} chunks={chunks} S={s_used:2d} ctas={tiles * chunks * s_used:3d}: {statistics.median(ts):6.1f} us   "
                  f"consumer reads {s_used * T * N * 4 / 1e6:5.1f} MB", flush=True)
capi.check(lib.blurr_set_global_option(b"op_gemm_bn", 0))
