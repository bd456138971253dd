#!/usr/bin/env python3
"""Per-CTA timeline of one GEMM launch (global option "gemm_cta_trace"): where a tile's time goes.
  python tools/cta_timeline.py [name-filter]
Stamps (gemm_body.cuh): 0 entry, 1 setup done, 2 producer past the dependency wait, 3 first stage landed,
4 last MMA issued, 5 accumulators complete, 6 TMEM drained into the smem tile, 7 tile stored."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from blurr_b200 import capi

lib = capi.load_library()
dev = torch.device("cuda:0")
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
SHAPES = [  # name, N, K, T, epi, splitk
    ("vlm gate/up", 32768, 2048, 276, capi.EPI_GEGLU, 1),
    ("vlm down", 2048, 16384, 276, capi.EPI_PARTIAL, 9),
    ("vlm qkv", 2560, 2048, 276, capi.EPI_PARTIAL, 7),
    ("vlm o", 2048, 2048, 276, capi.EPI_PARTIAL, 9),
    ("siglip qkv", 3456, 1152, 256, capi.EPI_STORE, 1),
    ("siglip fc1", 4352, 1152, 256, capi.EPI_GELU, 1),
    ("siglip fc2", 1152, 4352, 256, capi.EPI_PARTIAL, 16),
    ("siglip out", 1152, 1152, 256, capi.EPI_PARTIAL, 9),
]
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):      # e.g. OPTS="gemm_use_2cta=1,gemm_max_stages=6"
    k, v = kv.split("=")
    capi.check(lib.blurr_set_global_option(k.encode(), int(v)))
trace = torch.zeros((4096, 8), device=dev, dtype=torch.int64)
for name, N, K, T, epi, S in SHAPES:
    if ONLY and ONLY not in name:
        continue
    Ws = [torch.empty((N, K), device=dev, dtype=torch.bfloat16).uniform_(-0.02, 0.02) for _ in range(4)]
    X = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
    out = torch.empty((T, N), device=dev, dtype=torch.bfloat16)
    part = torch.empty((16 * T * N,), device=dev, dtype=torch.float32) if epi == capi.EPI_PARTIAL else None
    ldo = N // 2 if epi == capi.EPI_GEGLU else N

    def launch(i):
        return lib.blurr_op_gemm_async(sp, C.c_void_p(Ws[i % 4].data_ptr()), N, K, 0, C.c_void_p(X.data_ptr()), T, K,
                                       epi, S, None, C.c_void_p(out.data_ptr()), ldo,
                                       C.c_void_p(part.data_ptr()) if part is not None else None)
    for i in range(3):
        capi.check(launch(i))
    torch.cuda.synchronize()
    trace.zero_()
    capi.check(lib.blurr_set_global_option(b"gemm_cta_trace", trace.data_ptr()))
    capi.check(launch(3))
    torch.cuda.synchronize()
    capi.check(lib.blurr_set_global_option(b"gemm_cta_trace", 0))
    t = trace.cpu()
    used = t[:, 0] != 0
    t = t[used].double()
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    n = rel.shape[0]
    order = torch.argsort(rel[:, 0])
    print(f"== {name}: T={T} N={N} K={K} S={S}: {n} CTAs, kernel span {rel[:, 7].max():.2f} us")
    names = ["entry", "setup", "pdl", "first_full", "last_mma", "acc_ready", "drained", "stored"]
    for lab, sel in (("first wave (earliest 1/2)", order[: max(1, n // 2)]), ("last wave (latest 1/4)", order[-max(1, n // 4):])):
        r = rel[sel]
        print(f"  {lab}: " + "  ".join(f"{nm} {r[:, i].median():7.2f}" for i, nm in enumerate(names)))
        d = r[:, 1:] - r[:, :-1]
        print("    deltas (median): setup %.2f | pdl %.2f | first load %.2f | main loop %.2f | mma tail %.2f | drain %.2f | store %.2f"
              % tuple(d[:, i].median().item() for i in range(7)))
