#!/usr/bin/env python3
"""Batched-episode GEMM shapes (64 episodes: 16384 SigLIP tokens, 17664 prefix tokens) through the op-level C ABI,
back to back, against cuBLAS (torch.matmul) on the same shapes: which shapes lag the library ceiling.
  python tools/gemm_batched_bench.py [name-filter]            (OPTS="gemm_large_t_mode=1,..." for global options)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import capi

lib = capi.load_library()
dev = torch.device("cuda:0")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("=")
    capi.check(lib.blurr_set_global_option(k.encode(), int(v)))
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
SHAPES = [  # name, N, K, T, epilogue
    ("siglip patch", 1152, 640, 16384, capi.EPI_STORE),
    ("siglip qkv", 3456, 1152, 16384, capi.EPI_STORE),
    ("siglip out", 1152, 1152, 16384, capi.EPI_STORE),
    ("siglip fc1", 4352, 1152, 16384, capi.EPI_GELU),
    ("siglip fc2", 1152, 4352, 16384, capi.EPI_STORE),
    ("vlm qkv", 2560, 2048, 17664, capi.EPI_STORE),
    ("vlm o", 2048, 2048, 17664, capi.EPI_STORE),
    ("vlm gate/up", 32768, 2048, 17664, capi.EPI_GEGLU),
    ("vlm down", 2048, 16384, 17664, capi.EPI_STORE),
]
REPS = 10
for name, N, K, T, epi in SHAPES:
    if ONLY and ONLY not in name:
        continue
    W = torch.empty((N, K), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02)
    X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
    bias = torch.zeros((N,), device=dev, dtype=torch.bfloat16)
    ldo = N // 2 if epi == capi.EPI_GEGLU else N
    out = torch.empty((T, ldo), device=dev, dtype=torch.bfloat16)

    def ours():
        return lib.blurr_op_gemm_async(sp, C.c_void_p(W.data_ptr()), N, K, 0, C.c_void_p(X.data_ptr()), T, K, epi, 1,
                                       C.c_void_p(bias.data_ptr()) if epi != capi.EPI_GEGLU else None,
                                       C.c_void_p(out.data_ptr()), ldo, None)

    def cublas():
        return torch.matmul(X, W.t())

    res = {}
    for label, fn in (("ours", ours), ("cublas", cublas)):
        for _ in range(3):
            r = fn()
            if label == "ours":
                capi.check(min(r, 0))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(REPS):
            fn()
        b.record()
        torch.cuda.synchronize()
        res[label] = a.elapsed_time(b) * 1e3 / REPS
    fl = 2.0 * T * N * K
    print(f"{name:13s} T={T} N={N:5d} K={K:5d} epi{epi}: ours {res['ours']:7.1f} us = {fl / res['ours'] / 1e6:6.0f} TFLOP/s | "
          f"cuBLAS (no epilogue) {res['cublas']:7.1f} us = {fl / res['cublas'] / 1e6:6.0f} TFLOP/s | ratio {res['cublas'] / res['ours']:.2f}",
          flush=True)
