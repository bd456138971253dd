"""Per-kernel parity (-m gpu): each CUDA operator through its C-ABI entry point against a torch
restatement of the reference op sequence (same rounding points) on the same device."""

import math

import pytest
import torch

from blurr_b200 import capi
from helpers import bf16_ulp_err, op_gemm, op_joint_attention, op_siglip_attention, report

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


GEMM_SHAPES = [
    # (T, N, K)   the Pi-0 shapes of SURVEY.md appendix B (padded to the kernel's granularity)
    (256, 1152, 640), (256, 3456, 1152), (256, 4352, 1152), (256, 1152, 4352), (256, 2048, 1152),
    (276, 2560, 2048), (276, 2048, 2048), (276, 2048, 16384),
    (1, 2560, 1024), (4, 1024, 2048), (4, 1024, 4096), (4, 512, 1024),
    (552, 2560, 2048), (1104, 1152, 1152), (16, 128, 64), (17, 256, 192),
    (257, 512, 128), (288, 256, 192),      # edges of the two-tiles-per-CTA variant (256 < T <= 288)
    (2208, 1152, 1152), (1040, 256, 4352),  # > 1024 tokens: persistent tiles, double-buffered accumulator
]


@pytest.mark.parametrize("T,N,K", GEMM_SHAPES)
def test_gemm_store_bias(T, N, K):
    W = _rand((N, K), 1.0 / math.sqrt(K), 1)
    X = _rand((T, K), 1.0, 2)
    b = _rand((N,), 0.5, 3)
    got = op_gemm(W, X, capi.EPI_STORE, bias=b)
    ref = (X.float() @ W.float().t() + b.float()).to(torch.bfloat16)
    print(report(f"gemm_store T={T} N={N} K={K}", got, ref))
    assert bf16_ulp_err(got, ref) <= 1.01


@pytest.mark.parametrize("T,N,K,epi", [(552, 2560, 2048, "store"), (1104, 4352, 1152, "gelu"),
                                         (276, 4096, 2048, "geglu"), (16, 2560, 1024, "store"), (276, 2560, 2048, "store"),
                                         (270, 1152, 640, "gelu"), (288, 65536, 128, "geglu"),
                                         (2208, 2560, 2048, "store"), (2100, 4096, 1152, "geglu"), (4416, 1152, 640, "gelu"),
                                         (2304, 3456, 1152, "store")])      # odd tile counts: 9 and 27
def test_gemm_variants_agree(T, N, K, epi):
    """The GEMM variants (one CTA per tile, CTA pairs, persistent) accumulate every output in the same
    order, so they must agree bit for bit."""
    lib = capi.load_library()
    W = _rand((N, K), 1.0 / math.sqrt(K), 21)
    X = _rand((T, K), 1.0, 22)
    b = _rand((N,), 0.5, 23) if epi != "geglu" else None
    code = {"store": capi.EPI_STORE, "gelu": capi.EPI_GELU, "geglu": capi.EPI_GEGLU}[epi]
    outs = []
    try:
        for pairs, persistent, large, small in [(0, 0, 0, 0), (1, 0, 0, 0), (0, 1, 0, 0), (-1, 1, -1, 0), (0, 1, 1, 0), (-1, 1, 2, 0),
                                                (-1, 1, 3, 0), (-1, 1, -1, 2)]:
            capi.check(lib.blurr_set_global_option(b"gemm_use_2cta", pairs))
            capi.check(lib.blurr_set_global_option(b"gemm_persistent", persistent))
            capi.check(lib.blurr_set_global_option(b"gemm_large_t_mode", large))
            capi.check(lib.blurr_set_global_option(b"gemm_pair_small", small))      # persistent pairs for 257..288 tokens
            outs.append(op_gemm(W, X, code, bias=b))
    finally:
        capi.check(lib.blurr_set_global_option(b"gemm_use_2cta", -1))
        capi.check(lib.blurr_set_global_option(b"gemm_persistent", 1))
        capi.check(lib.blurr_set_global_option(b"gemm_large_t_mode", -1))
        capi.check(lib.blurr_set_global_option(b"gemm_pair_small", 1))
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("T,N,K,epi,band", [(2100, 4096, 1152, "geglu", 3), (2304, 3456, 1152, "store", 4),
                                              (5000, 4352, 640, "gelu", 5), (1500, 40960, 128, "geglu", 0)])
@pytest.mark.parametrize("policy", [0, 1, 2])
def test_gemm_pair_raster_bands(T, N, K, epi, band, policy):
    """Persistent CTA pairs visit the tiles in bands of weight tile pairs (last band narrower; band 0 = the
    automatic choice, which only bands when there are more weight tile pairs than CTA pairs: N = 40960).
    The raster order and the L2 eviction hints must not change a single bit."""
    lib = capi.load_library()
    W = _rand((N, K), 1.0 / math.sqrt(K), 41)
    X = _rand((T, K), 1.0, 42)
    b = _rand((N,), 0.5, 43) if epi != "geglu" else None
    code = {"store": capi.EPI_STORE, "gelu": capi.EPI_GELU, "geglu": capi.EPI_GEGLU}[epi]
    try:
        capi.check(lib.blurr_set_global_option(b"gemm_large_t_mode", 0))
        ref = op_gemm(W, X, code, bias=b)
        capi.check(lib.blurr_set_global_option(b"gemm_large_t_mode", -1))
        capi.check(lib.blurr_set_global_option(b"gemm_pair_band", band))
        capi.check(lib.blurr_set_global_option(b"gemm_pair_policy", policy))
        got = op_gemm(W, X, code, bias=b)
    finally:
        capi.check(lib.blurr_set_global_option(b"gemm_large_t_mode", -1))
        capi.check(lib.blurr_set_global_option(b"gemm_pair_band", 0))
        capi.check(lib.blurr_set_global_option(b"gemm_pair_policy", -1))
    assert torch.equal(got, ref)


@pytest.mark.parametrize("T,N,K", [(276, 2560, 2048), (17, 256, 192), (4, 1024, 4096)])
def test_gemm_row_major_weights(T, N, K):
    """Same kernel fed from a plain row-major nn.Linear weight (no packing)."""
    W = _rand((N, K), 1.0 / math.sqrt(K), 21)
    X = _rand((T, K), 1.0, 22)
    got = op_gemm(W, X, capi.EPI_STORE, packed=False)
    assert torch.equal(got, op_gemm(W, X, capi.EPI_STORE, packed=True))


@pytest.mark.parametrize("T,N,K,S", [(276, 2048, 16384, 9), (276, 2560, 2048, 7), (256, 1152, 4352, 16),
                                     (4, 1024, 4096, 16), (1, 2560, 1024, 6), (552, 2048, 2048, 3),
                                     (276, 2048, 2048, 1), (2208, 2048, 2048, 1), (1300, 2560, 640, 1)])
def test_gemm_partial_splitk(T, N, K, S):
    W = _rand((N, K), 1.0 / math.sqrt(K), 4)
    X = _rand((T, K), 1.0, 5)
    part = op_gemm(W, X, capi.EPI_PARTIAL, splitk=S)
    assert 1 <= part.shape[0] <= S
    got = part.sum(0)
    ref = X.float() @ W.float().t()
    err = (got - ref).abs().max().item()
    print(f"gemm_partial T={T} N={N} K={K} S={part.shape[0]}: max_abs={err:.3e}")
    assert err <= 2e-3


@pytest.mark.parametrize("T,N,K,S", [(276, 2048, 16384, 9), (276, 2560, 2048, 7), (260, 1152, 640, 3)])
def test_gemm_partial_splitk_pairs(T, N, K, S):
    """The persistent CTA-pair kernel with split-K slices as tiles (option gemm_pair_small=2) writes the same fp32
    partial sums as one CTA per (tile, slice): every output accumulates its k-blocks in the same order."""
    lib = capi.load_library()
    W = _rand((N, K), 1.0 / math.sqrt(K), 4)
    X = _rand((T, K), 1.0, 5)
    try:
        capi.check(lib.blurr_set_global_option(b"gemm_pair_small", 0))
        ref = op_gemm(W, X, capi.EPI_PARTIAL, splitk=S)
        capi.check(lib.blurr_set_global_option(b"gemm_pair_small", 2))
        got = op_gemm(W, X, capi.EPI_PARTIAL, splitk=S)
    finally:
        capi.check(lib.blurr_set_global_option(b"gemm_pair_small", 1))
    assert got.shape == ref.shape and torch.equal(got, ref)


@pytest.mark.parametrize("T,N,K", [(256, 4352, 1152), (17, 256, 192)])
def test_gemm_gelu(T, N, K):
    W = _rand((N, K), 1.0 / math.sqrt(K), 6)
    X = _rand((T, K), 1.0, 7)
    b = _rand((N,), 0.5, 8)
    got = op_gemm(W, X, capi.EPI_GELU, bias=b)
    h = (X.float() @ W.float().t() + b.float()).to(torch.bfloat16)
    ref = torch.nn.functional.gelu(h, approximate="tanh")
    print(report(f"gemm_gelu T={T} N={N} K={K}", got, ref))
    # a 1-ulp flip of the pre-activation (accumulation order) moves the output by <= 1 ulp of the
    # *input* magnitude; compare absolutely and bound the flip rate
    d = (got.float() - ref.float()).abs()
    assert d.max().item() <= 0.02 and (d > 0).float().mean().item() < 0.01


@pytest.mark.parametrize("T,I,K", [(276, 16384, 2048), (4, 4096, 1024), (17, 128, 192)])
def test_gemm_geglu(T, I, K):
    """Weight rows alternate gate_j, up_j (engine repack); output
    bf16(bf16(gelu(bf16(gate))) * bf16(up)) (paligemma/modules.py:93-95)."""
    Wg = _rand((I, K), 1.0 / math.sqrt(K), 9)
    Wu = _rand((I, K), 1.0 / math.sqrt(K), 10)
    X = _rand((T, K), 1.0, 11)
    Wi = torch.empty((2 * I, K), device=DEV, dtype=torch.bfloat16)
    Wi[0::2] = Wg
    Wi[1::2] = Wu
    got = op_gemm(Wi, X, capi.EPI_GEGLU)
    gate = (X.float() @ Wg.float().t()).to(torch.bfloat16)
    up = (X.float() @ Wu.float().t()).to(torch.bfloat16)
    ref = torch.nn.functional.gelu(gate, approximate="tanh") * up
    print(report(f"gemm_geglu T={T} I={I} K={K}", got, ref))
    d = (got.float() - ref.float()).abs()
    assert d.max().item() <= 0.05 and (d > 0).float().mean().item() < 0.02


@pytest.mark.parametrize("batch", [1, 3, 6])      # 16-, 32- and 64-row tiles
def test_siglip_attention(batch):
    """siglip.py:133-152: bf16 QK^T * scale, fp32 softmax -> bf16, PV."""
    seq, heads, hidden = 256, 16, 1152
    hd = hidden // heads
    qkv = _rand((batch * seq, 3 * hidden), 1.0, 12)
    got = op_siglip_attention(qkv, batch, seq, heads, hidden)
    q, k, v = [t.view(batch, seq, heads, hd).transpose(1, 2) for t in qkv.view(batch, seq, 3 * hidden).split(hidden, -1)]
    w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(torch.bfloat16)
    ref = torch.matmul(w, v).transpose(1, 2).contiguous().view(batch * seq, hidden)
    print(report(f"siglip_attention B={batch}", got, ref))
    assert (got.float() - ref.float()).abs().max().item() <= 0.03


@pytest.mark.parametrize("batch,seq,scale", [(1, 256, 1.0), (2, 256, 3.0), (1, 200, 1.0)])
def test_siglip_attention_stream(batch, seq, scale):
    """SigLIP attention of one or two images as the streaming kernel (32 query rows of one head per CTA, K straight into
    permuted mma.sync fragments, V in shared memory) against the torch restatement and the mma.sync tile kernel."""
    lib = capi.load_library()
    heads, hidden = 16, 1152
    hd = hidden // heads
    qkv = _rand((batch * seq, 3 * hidden), scale, 62)
    try:
        capi.check(lib.blurr_set_global_option(b"attn_siglip_stream", 0))
        capi.check(lib.blurr_set_global_option(b"attn_tc", 0))
        base = op_siglip_attention(qkv, batch, seq, heads, hidden)
        capi.check(lib.blurr_set_global_option(b"attn_siglip_stream", 1))
        got = op_siglip_attention(qkv, batch, seq, heads, hidden)
    finally:
        capi.check(lib.blurr_set_global_option(b"attn_tc", -1))
        capi.check(lib.blurr_set_global_option(b"attn_siglip_stream", 1))
    q, k, v = [t.view(batch, seq, heads, hd).transpose(1, 2) for t in qkv.view(batch, seq, 3 * hidden).split(hidden, -1)]
    w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(torch.bfloat16)
    ref = torch.matmul(w, v).transpose(1, 2).contiguous().view(batch * seq, hidden)
    print(report(f"siglip_attention stream B={batch} seq={seq} scale={scale}", got, ref))
    print(report("  vs mma.sync kernel", got, base))
    assert (got.float() - ref.float()).abs().max().item() <= 0.03 * scale
    d = (got.float() - base.float()).abs()
    assert d.max().item() <= 0.03 * scale and (d > 0).float().mean().item() < 0.05


@pytest.mark.parametrize("batch", [1, 3, 9])
def test_siglip_attention_tcgen05(batch):
    """Batched-episode SigLIP attention on tcgen05 (head_dim 72 zero-padded to 128 inside the Q tile) against
    the torch restatement and the mma.sync kernel."""
    lib = capi.load_library()
    seq, heads, hidden = 256, 16, 1152
    hd = hidden // heads
    qkv = _rand((batch * seq, 3 * hidden), 1.0, 52)
    try:
        capi.check(lib.blurr_set_global_option(b"attn_siglip_stream", 0))
        capi.check(lib.blurr_set_global_option(b"attn_tc", 0))
        base = op_siglip_attention(qkv, batch, seq, heads, hidden)
        capi.check(lib.blurr_set_global_option(b"attn_tc", 1))
        got = op_siglip_attention(qkv, batch, seq, heads, hidden)
    finally:
        capi.check(lib.blurr_set_global_option(b"attn_tc", -1))
        capi.check(lib.blurr_set_global_option(b"attn_siglip_stream", 1))
    q, k, v = [t.view(batch, seq, heads, hd).transpose(1, 2) for t in qkv.view(batch, seq, 3 * hidden).split(hidden, -1)]
    w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(torch.bfloat16)
    ref = torch.matmul(w, v).transpose(1, 2).contiguous().view(batch * seq, hidden)
    print(report(f"siglip_attention tcgen05 B={batch}", got, ref))
    print(report("  vs mma.sync kernel", got, base))
    assert (got.float() - ref.float()).abs().max().item() <= 0.03
    assert (got.float() - base.float()).abs().max().item() <= 0.03
    assert ((got.float() - base.float()).abs() > 0).float().mean().item() < 0.05


def _joint_ref(q, kc, vc, mask_rows, n_heads):
    """joint_model.py:273-288 on [B, H, Q, 256] / [B, 1, KV, 256]."""
    B = kc.shape[0]
    qh = q.view(B, -1, n_heads, 256).transpose(1, 2)
    k = kc[:, None].expand(B, n_heads, kc.shape[1], 256)
    v = vc[:, None].expand(B, n_heads, vc.shape[1], 256)
    w = torch.matmul(qh, k.transpose(2, 3)) / math.sqrt(256)
    w = w / 50.0
    w = torch.tanh(w)
    w = w * 50.0
    w = w + mask_rows[:, None]
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v)
    return o.transpose(1, 2).contiguous().view(q.shape)


def _block_mask(B, rows, cols, cnts, row0):
    m = torch.full((B, rows, cols), torch.finfo(torch.bfloat16).min, dtype=torch.bfloat16, device=DEV)
    for b, c in enumerate(cnts):
        for r in range(rows):
            gr = row0 + r
            if gr < 276:
                if gr < c:
                    m[b, r, :c] = 0
            else:
                m[b, r, :c] = 0
                m[b, r, 276:min(cols, 277)] = 0
                if gr >= 277:
                    m[b, r, 276:cols] = 0
    return m


@pytest.mark.parametrize("batch,scale", [(1, 1.0), (2, 6.0), (4, 2.0)])     # batch 4: 32-row tiles
def test_joint_attention_prefill(batch, scale):
    n_heads, n_keys, slots = 8, 277, 281
    q = _rand((batch * 276, n_heads * 256), scale, 13)
    kc = _rand((batch, slots, 256), scale, 14)
    vc = _rand((batch, slots, 256), 1.0, 15)
    mask = _block_mask(batch, 277, 277, [268, 261, 276, 1][:batch], 0)
    got = op_joint_attention(False, q, 276, 0, kc, vc, n_keys, mask, batch, n_heads)
    ref = _joint_ref(q, kc[:, :n_keys], vc[:, :n_keys], mask[:, :276], n_heads)
    print(report(f"joint_prefill B={batch} scale={scale}", got, ref))
    assert (got.float() - ref.float()).abs().max().item() <= 0.05


@pytest.mark.parametrize("batch,scale", [(1, 1.0), (2, 6.0), (9, 2.0)])
def test_joint_attention_prefill_tcgen05(batch, scale):
    """The batched-episode variant (tcgen05 S = QK^T and O = PV, V as an MN-major operand, rows = (head, query)
    pairs) against the torch restatement and against the mma.sync kernel.  The two kernels round at the same
    points; only fp32 summation orders differ."""
    lib = capi.load_library()
    n_heads, n_keys, slots = 8, 277, 281
    q = _rand((batch * 276, n_heads * 256), scale, 43)
    kc = _rand((batch, slots, 256), scale, 44)
    vc = _rand((batch, slots, 256), 1.0, 45)
    cnts = ([268, 261, 276, 258] * 3)[:batch]
    mask = _block_mask(batch, 277, 277, cnts, 0)
    try:
        capi.check(lib.blurr_set_global_option(b"attn_prefill_stream", 0))
        capi.check(lib.blurr_set_global_option(b"attn_tc", 0))
        base = op_joint_attention(False, q, 276, 0, kc, vc, n_keys, mask, batch, n_heads)
        capi.check(lib.blurr_set_global_option(b"attn_tc", 1))
        got = op_joint_attention(False, q, 276, 0, kc, vc, n_keys, mask, batch, n_heads)
    finally:
        capi.check(lib.blurr_set_global_option(b"attn_tc", -1))
        capi.check(lib.blurr_set_global_option(b"attn_prefill_stream", 1))
    ref = _joint_ref(q, kc[:, :n_keys], vc[:, :n_keys], mask[:, :276], n_heads)
    print(report(f"joint_prefill tcgen05 B={batch} scale={scale}", got, ref))
    print(report("  vs mma.sync kernel", got, base))
    assert (got.float() - ref.float()).abs().max().item() <= 0.05
    assert (got.float() - base.float()).abs().max().item() <= 0.05
    assert ((got.float() - base.float()).abs() > 0).float().mean().item() < 0.05


@pytest.mark.parametrize("batch,scale,qps", [(1, 1.0, 276), (2, 2.0, 276), (2, 6.0, 276), (1, 3.0, 37)])
def test_joint_attention_prefill_stream(batch, scale, qps):
    """The one-or-two-episode prefill kernel (16 rows = 2 queries x 8 heads per CTA, K straight into permuted mma.sync
    fragments, V in shared memory) against the torch restatement and the mma.sync tile kernel: same rounding points,
    the 256 dims of a dot product are only summed in another order."""
    lib = capi.load_library()
    n_heads, n_keys, slots = 8, 277, 281
    q = _rand((batch * qps, n_heads * 256), scale, 53)
    kc = _rand((batch, slots, 256), scale, 54)
    vc = _rand((batch, slots, 256), 1.0, 55)
    mask = _block_mask(batch, 277, 277, [268, 261][:batch], 0)
    try:
        capi.check(lib.blurr_set_global_option(b"attn_prefill_stream", 0))
        capi.check(lib.blurr_set_global_option(b"attn_tc", 0))
        base = op_joint_attention(False, q, qps, 0, kc, vc, n_keys, mask, batch, n_heads)
        capi.check(lib.blurr_set_global_option(b"attn_prefill_stream", 1))
        got = op_joint_attention(False, q, qps, 0, kc, vc, n_keys, mask, batch, n_heads)
    finally:
        capi.check(lib.blurr_set_global_option(b"attn_tc", -1))
        capi.check(lib.blurr_set_global_option(b"attn_prefill_stream", 1))
    ref = _joint_ref(q, kc[:, :n_keys], vc[:, :n_keys], mask[:, :qps], n_heads)
    print(report(f"joint_prefill stream B={batch} scale={scale} qps={qps}", got, ref))
    print(report("  vs mma.sync kernel", got, base))
    assert (got.float() - ref.float()).abs().max().item() <= 0.05
    # against the other kernel: a logit that lands on a bf16 rounding boundary may round the other way (the dims are
    # summed in another order); with peaked softmaxes (scale 6) one such flip moves an output row visibly, so the bound
    # there is on how many elements differ at all, not on the largest one
    d = (got.float() - base.float()).abs()
    assert (d > 0).float().mean().item() < 0.01 and d.mean().item() < 1e-4
    if scale <= 3.0:
        assert d.max().item() <= 0.05


@pytest.mark.parametrize("qps,row0,n_keys,batch", [(1, 276, 277, 2), (4, 0, 281, 2), (4, 0, 281, 160)])
def test_joint_attention_fewq(qps, row0, n_keys, batch):
    n_heads, slots = 8, 281
    cnts = [268, 270] + [200 + (i * 7) % 77 for i in range(batch - 2)]
    q = _rand((batch * qps, n_heads * 256), 4.0, 16)
    kc = _rand((batch, slots, 256), 4.0, 17)
    vc = _rand((batch, slots, 256), 1.0, 18)
    if qps == 1:
        mask = _block_mask(batch, 277, 277, cnts, 0)      # proprio row = row 276 of the prefill mask
        rows = mask[:, 276:277]
    else:
        mask = _block_mask(batch, 4, 281, cnts, 277)      # action mask rows
        rows = mask
    lib = capi.load_library()
    ref = _joint_ref(q, kc[:, :n_keys], vc[:, :n_keys], rows, n_heads)
    outs = {}
    try:
        # "stream": the default streaming kernel (K rows in registers, V in shared memory, FMA dot products);
        # 0: mma.sync tile kernel; 1: the tcgen05 kernel (one 128-row tile of (head, query) pairs per sample)
        for mode in ("stream", 0, 1):
            capi.check(lib.blurr_set_global_option(b"attn_fewq_stream", 1 if mode == "stream" else 0))
            capi.check(lib.blurr_set_global_option(b"attn_tc_fewq", mode if mode != "stream" else 0))
            outs[mode] = op_joint_attention(True, q, qps, row0, kc, vc, n_keys, mask, batch, n_heads)
            print(report(f"joint_fewq qps={qps} kernel={mode}", outs[mode], ref))
            assert (outs[mode].float() - ref.float()).abs().max().item() <= 0.05
    finally:
        capi.check(lib.blurr_set_global_option(b"attn_tc_fewq", -1))
        capi.check(lib.blurr_set_global_option(b"attn_fewq_stream", 1))
    ds = (outs["stream"].float() - outs[0].float()).abs()
    assert ds.max().item() <= 0.05 and (ds > 0).float().mean().item() < 0.05
    d = (outs[0].float() - outs[1].float()).abs()
    assert d.max().item() <= 0.05 and (d > 0).float().mean().item() < 0.05
