#!/usr/bin/env python3
"""Run the dominant kernel (Gemma gate/up GEMM + GeGLU, 276 tokens) a few times on rotating
weight buffers (tile-packed layout, ldw = 0, as the engine streams them), for `ncu --set full` captures:  ncu ... -k regex:gemm_tc_kernel python tools/prof_gemm.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import capi

which = sys.argv[1] if len(sys.argv) > 1 else "gateup"
lib = capi.load_library()
dev = torch.device("cuda:0")
shapes = {"gateup": (32768, 2048, 276, capi.EPI_GEGLU, 1), "down": (2048, 16384, 276, capi.EPI_PARTIAL, 9),
          "qkv": (2560, 2048, 276, capi.EPI_PARTIAL, 7), "fc1": (4352, 1152, 256, capi.EPI_GELU, 1),
          "gateup64": (32768, 2048, 17664, capi.EPI_GEGLU, 1)}
N, K, T, epi, S = shapes[which]
nbuf = 4
Ws = [torch.empty((N, K), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02) for _ in range(nbuf)]
X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
out = torch.empty((T, N), device=dev, dtype=torch.bfloat16)
part = torch.empty((16, T, N), device=dev, dtype=torch.float32) if epi == capi.EPI_PARTIAL else None
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for i in range(8):
    capi.check(lib.blurr_op_gemm(sp, C.c_void_p(Ws[i % nbuf].data_ptr()), N, K, 0, C.c_void_p(X.data_ptr()), T, K, epi,
                                 S, None, C.c_void_p(out.data_ptr()), N // 2 if epi == capi.EPI_GEGLU else N,
                                 C.c_void_p(part.data_ptr()) if part is not None else None))
torch.cuda.synchronize()
print("ok")
