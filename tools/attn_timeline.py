#!/usr/bin/env python3
"""Per-CTA timeline of the tcgen05 joint-attention kernel (global option "attn_cta_trace"), few-query and prefill shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from blurr_b200 import capi
from helpers import op_joint_attention

lib = capi.load_library()
dev = "cuda"
n_heads, slots = 8, 281
trace = torch.zeros((4096, 8), device=dev, dtype=torch.int64)
names = ["entry", "pass2_done", "q_staged", "s_done", "p_written", "pass1_done", "o_done", "stored"]
for label, fewq, qps, n_keys, batch in [("action few-query (4 q)", True, 4, 281, 1), ("proprio few-query (1 q)", True, 1, 277, 1),
                                       ("prefill (276 q)", False, 276, 277, 1)]:
    q = torch.randn((batch * qps, n_heads * 256), device=dev).to(torch.bfloat16)
    kc = torch.randn((batch, slots, 256), device=dev).to(torch.bfloat16)
    vc = torch.randn((batch, slots, 256), device=dev).to(torch.bfloat16)
    rows = qps if fewq and qps == 4 else 277
    cols = 281 if qps == 4 and fewq else 277
    # rows padded to a multiple of 8 elements like the engine's staged masks (16-byte mask loads in the kernel)
    mask = torch.zeros((batch, rows, (cols + 7) // 8 * 8), device=dev, dtype=torch.bfloat16)[:, :, :cols]
    capi.check(lib.blurr_set_global_option(b"attn_tc", 1))
    capi.check(lib.blurr_set_global_option(b"attn_tc_fewq", 1))
    row0 = 276 if (fewq and qps == 1) else 0
    for _ in range(3):
        op_joint_attention(fewq, q, qps, row0, kc, vc, n_keys, mask, batch, n_heads)
    trace.zero_()
    capi.check(lib.blurr_set_global_option(b"attn_cta_trace", trace.data_ptr()))
    op_joint_attention(fewq, q, qps, row0, kc, vc, n_keys, mask, batch, n_heads)
    capi.check(lib.blurr_set_global_option(b"attn_cta_trace", 0))
    t = trace.cpu()
    t = t[t[:, 0] != 0].double()
    rel = (t - t[:, 0].min()) / 1e3
    print(f"== {label}: {rel.shape[0]} CTAs, span {rel.max():.2f} us")
    print("   " + "  ".join(f"{n} {rel[:, i][rel[:, i] >= 0].median():6.2f}" for i, n in enumerate(names)))
capi.check(lib.blurr_set_global_option(b"attn_tc", -1))
capi.check(lib.blurr_set_global_option(b"attn_tc_fewq", -1))
