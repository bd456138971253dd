#!/usr/bin/env python3
"""Time the GEMM kernel on the Pi-0 shapes (CUDA events, rotating weight buffers > L2)."""
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import capi

lib = capi.load_library()
dev = torch.device("cuda:0")
ONLY = os.environ.get("ONLY", "")
SHAPES = [  # name, N, K, T, epi, splitk
    ("vlm gate/up", 32768, 2048, 276, capi.EPI_GEGLU, 1),
    ("vlm down", 2048, 16384, 276, capi.EPI_PARTIAL, 9),
    ("vlm qkv", 2560, 2048, 276, capi.EPI_PARTIAL, 7),
    ("vlm o", 2048, 2048, 276, capi.EPI_PARTIAL, 9),
    ("siglip qkv", 3456, 1152, 256, capi.EPI_STORE, 1),
    ("siglip fc1", 4352, 1152, 256, capi.EPI_GELU, 1),
    ("siglip fc2", 1152, 4352, 256, capi.EPI_PARTIAL, 16),
    ("expert gate/up", 8192, 1024, 4, capi.EPI_GEGLU, 1),
    ("expert down", 1024, 4096, 4, capi.EPI_PARTIAL, 16),
    ("vlm gate/up bs64", 32768, 2048, 17664, capi.EPI_GEGLU, 1),
    ("vlm down bs64", 2048, 16384, 17664, capi.EPI_PARTIAL, 1),
    ("siglip fc1 bs64", 4352, 1152, 16384, capi.EPI_GELU, 1),
    ("siglip qkv bs64", 3456, 1152, 16384, capi.EPI_STORE, 1),
    ("vlm qkv bs64", 2560, 2048, 17664, capi.EPI_PARTIAL, 1),
    ("vlm down bs64 store", 2048, 16384, 17664, capi.EPI_STORE, 1),
    ("vlm qkv bs64 store", 2560, 2048, 17664, capi.EPI_STORE, 1),
    ("siglip fc2 bs64 store", 1152, 4352, 16384, capi.EPI_STORE, 1),
]
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
# weights: tile-packed (ldw = 0; values are random so no packing pass is needed) vs row-major (ldw = K)
CASES = [(int(a), "packed") for a in (sys.argv[1:] or ["1", "0"])]
for use2, LDW_MODE in CASES:
    cmax = use2
    capi.check(lib.blurr_set_global_option(b"gemm_use_2cta", use2))
    capi.check(lib.blurr_set_global_option(b"gemm_large_t_mode", int(os.environ.get("MODE", "0"))))
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):      # e.g. OPTS="gemm_pair_band=16,gemm_pair_policy=1"
        k, v = kv.split("=")
        capi.check(lib.blurr_set_global_option(k.encode(), int(v)))
    for name, N, K, T, epi, S in SHAPES:
        if ONLY and ONLY not in name:
            continue
        LDW = 0 if LDW_MODE == "packed" else K
        nbuf = max(2, min(6, int(600e6 // (N * K * 2)) + 1))
        Ws = [torch.empty((N, K), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02) for _ in range(nbuf)]
        X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
        out = torch.empty((T, N), device=dev, dtype=torch.bfloat16)
        part = torch.empty((16 * T * N if epi == capi.EPI_PARTIAL and T < 1000 else T * N,), device=dev, dtype=torch.float32) \
            if epi == capi.EPI_PARTIAL else None
        ldo = N // 2 if epi == capi.EPI_GEGLU else N

        def launch(i):
            return lib.blurr_op_gemm_async(sp, C.c_void_p(Ws[i % nbuf].data_ptr()), N, K, LDW, C.c_void_p(X.data_ptr()), T, K,
                                           epi, S, None, C.c_void_p(out.data_ptr()), ldo,
                                           C.c_void_p(part.data_ptr()) if part is not None else None)
        for i in range(3):
            capi.check(launch(i))
        torch.cuda.synchronize()
        iters = 20 if T < 1000 else 5
        pairs = []
        for i in range(iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); capi.check(launch(i)); e.record()
            pairs.append((s, e))
        torch.cuda.synchronize()
        ms = statistics.fmean(s.elapsed_time(e) for s, e in pairs)
        gbs = N * K * 2 / ms / 1e6
        tf = 2.0 * N * K * T / ms / 1e9
        print(f"2cta={cmax} {LDW_MODE:8s} {name:18s} T={T:5d} N={N:5d} K={K:5d}: {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s weights  {tf:7.1f} TFLOP/s", flush=True)
        del Ws, X, out, part
