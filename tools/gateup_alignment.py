#!/usr/bin/env python3
"""Does the stand-alone gate/up GEMM time depend on where its weight buffers sit?  Back-to-back launches over 4 weight
buffers at different spacings (bench.py's roofline leg was bimodal from run to run: 38 vs 47 us)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import capi

lib = capi.load_library()
dev = torch.device("cuda:0")
N, K, T = 32768, 2048, 276
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
out = torch.empty((T, N // 2), device=dev, dtype=torch.bfloat16)
elems = N * K


def run(name, ptrs, iters=40):
    def launch(i):
        return lib.blurr_op_gemm_async(sp, C.c_void_p(ptrs[i % len(ptrs)]), N, K, 0, C.c_void_p(X.data_ptr()), T, K,
                                       capi.EPI_GEGLU, 1, None, C.c_void_p(out.data_ptr()), N // 2, None)
    for i in range(len(ptrs)):
        capi.check(launch(i))
    torch.cuda.synchronize()
    res = []
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(iters):
            launch(i)
        e.record()
        torch.cuda.synchronize()
        res.append(s.elapsed_time(e) / iters * 1e3)
    print(f"{name:44s}: " + "  ".join(f"{r:6.2f}" for r in res) + " us per launch", flush=True)


for pad_kb in (0, 4, 64, 256, 1024, 1536, 2048 + 64):
    pad = pad_kb * 1024 // 2
    big = torch.empty((4 * (elems + pad) + 2 ** 21,), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02)
    base = big.data_ptr()
    run(f"one block, spacing 128 MB + {pad_kb} KB", [base + i * (elems + pad) * 2 for i in range(4)])
    del big
    torch.cuda.empty_cache()
for off_kb in (0, 64, 512, 1024):
    big = torch.empty((elems + 2 ** 21,), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02)
    run(f"single buffer re-used, base offset {off_kb} KB", [big.data_ptr() + off_kb * 1024])
    del big
    torch.cuda.empty_cache()
junk = [torch.empty((37 * 2 ** 20 + 12345,), device=dev, dtype=torch.bfloat16) for _ in range(7)]
sep = [torch.empty((elems,), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02) for _ in range(4)]
run("4 separate allocations (after junk)", [t.data_ptr() for t in sep])
print("bases:", [hex(t.data_ptr()) for t in sep])
