#!/usr/bin/env python3
"""Ring-depth sensitivity of the 1-CTA and CTA-pair GEMM on the gate/up and down shapes (276 tokens)."""
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
SHAPES = [("gate/up", 32768, 2048, capi.EPI_GEGLU, 1), ("down", 2048, 16384, capi.EPI_PARTIAL, 9)]
if os.environ.get("EXPERT"):
    SHAPES = [("x qkv", 2560, 1024, capi.EPI_PARTIAL, 7), ("x o", 1024, 2048, capi.EPI_PARTIAL, 16),
              ("x gate/up", 8192, 1024, capi.EPI_GEGLU, 1), ("x down", 1024, 4096, capi.EPI_PARTIAL, 16)]
TOKENS = [int(a) for a in sys.argv[1:]] or [276]
PAIRS = [int(a) for a in os.environ.get('PAIRS', '0,1').split(',')]
PERSIST = [int(a) for a in os.environ.get('PERSIST', '1').split(',')]
STAGES = [int(a) for a in os.environ.get('STAGES', '2,3,4,6,8').split(',')]
for name, N, K, epi, S in SHAPES:
    nbuf = 5
    Ws = [torch.empty((N, K), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02) for _ in range(nbuf)]
    for T in TOKENS:
        X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
        out = torch.empty((T, N), device=dev, dtype=torch.bfloat16)
        part = torch.empty((16 * T * N,), device=dev, dtype=torch.float32) if epi == capi.EPI_PARTIAL else None
        ldo = N // 2 if epi == capi.EPI_GEGLU else N
        for use2 in PAIRS:
            for stages, persist in [(a, b) for a in STAGES for b in PERSIST]:
                capi.check(lib.blurr_set_global_option(b"gemm_persistent", persist))
                capi.check(lib.blurr_set_global_option(b"gemm_use_2cta", use2))
                capi.check(lib.blurr_set_global_option(b"gemm_max_stages", stages))

                def launch(i):
                    return lib.blurr_op_gemm_async(sp, C.c_void_p(Ws[i % nbuf].data_ptr()), N, K, 0, C.c_void_p(X.data_ptr()),
                                                   T, K, epi, S, None, C.c_void_p(out.data_ptr()), ldo,
                                                   C.c_void_p(part.data_ptr()) if part is not None else None)
                for i in range(3):
                    capi.check(launch(i))
                torch.cuda.synchronize()
                pairs = []
                for i in range(20):
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); capi.check(launch(i)); e.record()
                    pairs.append((s, e))
                torch.cuda.synchronize()
                ms = statistics.fmean(s.elapsed_time(e) for s, e in pairs)
                print(f"{name:8s} T={T:4d} pairs={use2} persistent={persist} max_stages={stages}: {ms * 1e3:7.1f} us  {N * K * 2 / ms / 1e6:6.0f} GB/s", flush=True)
