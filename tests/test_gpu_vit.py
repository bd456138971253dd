"""Vision side of the OpenVLA-7B-shaped path (include/blurr_vit.h): the generic ViT tower against transformers'
`Dinov2WithRegistersModel` (cls + register tokens, LayerScale, exact GELU) and `SiglipVisionModel` (tanh GELU, an MLP width
that is not a multiple of 128) on the same bf16 weights - patch tokens of the second-to-last block, the features
OpenVLA's fused backbone concatenates - and the 3-layer GELU projector against torch.  Like the language model, parity is
pinned to transformers 5.5 (a library of this image); the reference's own remote code is not available."""

import pytest
import torch

from blurr_b200 import openvla

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _report(name, got, ref):
    d = (got.float() - ref.float()).abs()
    rms = ref.float().pow(2).mean().sqrt().item()
    print(f"{name}: max_abs {d.max().item():.3e} mean_abs {d.mean().item():.3e} ref_rms {rms:.3e}")
    return d.max().item(), d.mean().item(), rms


def _check_against_fp32(name, got, ref16, ref32):
    """bf16 towers legitimately drift apart (different GEMM summation orders, amplified by every LayerNorm): the fp32
    run of the same bf16 weights is the tie-breaker - we must be as close to it as transformers' own bf16 run is."""
    e_ours = (got.float() - ref32).abs()
    e_ref = (ref16.float() - ref32).abs()
    rms = ref32.pow(2).mean().sqrt().item()
    print(f"{name}: vs fp32 ours max {e_ours.max().item():.3e} mean {e_ours.mean().item():.3e} | transformers bf16 max "
          f"{e_ref.max().item():.3e} mean {e_ref.mean().item():.3e} | ours vs transformers bf16 max "
          f"{(got.float() - ref16.float()).abs().max().item():.3e} (ref rms {rms:.3f})")
    assert e_ours.mean().item() <= 1.5 * e_ref.mean().item() + 1e-4 * rms
    assert e_ours.max().item() <= 2.5 * e_ref.max().item() + 1e-2 * rms


def _pixels(batch, seed):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    return (torch.rand((batch, 3, 224, 224), device=DEV, generator=g) * 2 - 1).to(torch.bfloat16)


# (hidden, heads, depth, batch): small towers, and DINOv2-L's real width at 5 images (1305 rows: the bf16 hand-off GEMMs and
# the streaming consumers, with the prefix-row map and LayerScale)
@pytest.mark.parametrize("hidden,heads,depth,batch", [(256, 4, 4, 1), (256, 4, 4, 2), (1024, 16, 3, 5)])
def test_dinov2_registers_tower_matches_transformers(hidden, heads, depth, batch):
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    hc = Dinov2WithRegistersConfig(hidden_size=hidden, num_hidden_layers=depth, num_attention_heads=heads, mlp_ratio=4, image_size=224,
                                   patch_size=14, num_register_tokens=4, layerscale_value=1.0, attn_implementation="eager")
    torch.manual_seed(0)
    model = Dinov2WithRegistersModel(hc).eval()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "lambda1" in name:
                p.copy_(0.5 + torch.rand_like(p))                     # LayerScale away from 1
            elif "cls_token" in name or "register_tokens" in name or "position_embeddings" in name:
                p.copy_(0.5 * torch.randn_like(p))
            elif name.endswith("weight") and p.dim() == 2:
                p.mul_(3.0 if hidden <= 256 else 1.5)                 # attention and the MLP actually move the stream
    model = model.to(torch.bfloat16).to(DEV)
    enc = openvla.VitEncoder.from_hf_dinov2(model, DEV, max_batch=batch)
    px = _pixels(batch, 1)
    with torch.inference_mode():
        ref = model(pixel_values=px, output_hidden_states=True).hidden_states[-2][:, 5:]
        ref32 = model.float()(pixel_values=px.float(), output_hidden_states=True).hidden_states[-2][:, 5:]
    got = enc.forward(px)
    torch.cuda.synchronize()
    _check_against_fp32(f"dinov2 tower B={batch} ({enc.last_launch_count} launches)", got, ref, ref32)
    again = enc.forward(px)                        # CUDA-graph replay
    enc.set_option("use_cuda_graph", 0)
    eager = enc.forward(px)
    torch.cuda.synchronize()
    assert torch.equal(got, again) and torch.equal(got, eager)
    enc.close()


# small towers, and SigLIP-so400m's real widths (1152 / 4304, head_dim 72) at 5 images
@pytest.mark.parametrize("hidden,inter,heads,depth,batch", [(256, 560, 4, 3, 1), (256, 560, 4, 3, 3), (1152, 4304, 16, 3, 5)])
def test_siglip_tower_matches_transformers(hidden, inter, heads, depth, batch):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    hc = SiglipVisionConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=depth, num_attention_heads=heads, image_size=224,
                            patch_size=14, attn_implementation="eager")
    torch.manual_seed(1)
    model = SiglipVisionModel(hc).eval()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("weight") and p.dim() == 2 and "position" not in name:
                p.mul_(2.0 if hidden <= 256 else 1.5)
    model = model.to(torch.bfloat16).to(DEV)
    enc = openvla.VitEncoder.from_hf_siglip(model, DEV, max_batch=batch)
    px = _pixels(batch, 2)
    with torch.inference_mode():
        ref = model(pixel_values=px, output_hidden_states=True).hidden_states[-2]
        ref32 = model.float()(pixel_values=px.float(), output_hidden_states=True).hidden_states[-2]
    got = enc.forward(px)
    torch.cuda.synchronize()
    _check_against_fp32(f"siglip tower B={batch} ({enc.last_launch_count} launches)", got, ref, ref32)
    enc.close()


def test_projector_and_fused_backbone():
    dims = [2176, 8704, 4096, 4096]
    torch.manual_seed(2)
    ref_mlp = torch.nn.Sequential(torch.nn.Linear(dims[0], dims[1]), torch.nn.GELU(), torch.nn.Linear(dims[1], dims[2]),
                                  torch.nn.GELU(), torch.nn.Linear(dims[2], dims[3])).to(torch.bfloat16).to(DEV)
    proj = openvla.MlpProjector(dims, DEV, max_rows=512)
    for i, m in enumerate([ref_mlp[0], ref_mlp[2], ref_mlp[4]]):
        proj.set_layer(i, m.weight, m.bias)
    x = torch.randn((2, 256, dims[0]), device=DEV).to(torch.bfloat16)
    with torch.inference_mode():
        ref = ref_mlp(x)
    got = proj.forward(x)
    torch.cuda.synchronize()
    mx, mean, rms = _report("projector 2176-8704-4096-4096", got, ref)
    assert mx <= 0.05 * rms + 0.01 and mean <= 0.005 * rms
    # fused backbone: both towers fill one [B, 256, 2176] feature matrix, DINOv2 columns first
    dino = openvla.VitEncoder.synthetic(openvla.dinov2_large_reg4_config(depth=3), DEV, max_batch=2, seed=3)
    sig = openvla.VitEncoder.synthetic(openvla.siglip_so400m_config(depth=3), DEV, max_batch=2, seed=4)
    fused = openvla.FusedVisionBackbone(dino, sig, proj)
    px = _pixels(2, 5)
    out = fused.forward(px, px)
    torch.cuda.synchronize()
    feats = torch.cat([dino.forward(px), sig.forward(px)], dim=-1)
    assert out.shape == (2, 256, 4096) and torch.isfinite(out.float()).all()
    assert torch.equal(out, proj.forward(feats))
    serial = openvla.FusedVisionBackbone(dino, sig, proj, concurrent=False).forward(px, px)
    torch.cuda.synchronize()
    assert torch.equal(out, serial)                  # two streams or one: same bits
    for m in (dino, sig, proj):
        m.close()
