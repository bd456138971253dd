"""Full-size Pi-0 Bridge (27 SigLIP + 18 joint layers, 3.55 B parameters) on the GPU (-m gpu):
the CUDA control step against the oracle's bf16 run on the same device (BASELINE.json north_star:
max-abs action error <= 1e-2 on the clamped output, per-layer errors reported), the fp32 tie-break,
episode-sharding invariance and the KV-cache slot layout."""

import pytest
import torch

from blurr_b200 import dist as bdist
from blurr_b200 import synth
from blurr_b200.config import bridge_config
from blurr_b200.pizero import PiZeroInference
from oracle import pi0_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def full():
    cfg = bridge_config(1)
    cfg.final_action_clip_value = None
    sd = synth.synthetic_state_dict(cfg, 0, torch.bfloat16)
    model = PiZeroInference.from_state_dict(cfg, sd, device=DEV)
    model.set_engine_options(reserve_batch=4)
    sd_gpu = {k: v.to(DEV) for k, v in sd.items()}
    yield cfg, model, sd_gpu
    model.release_engine()


def _oracle(sd, cfg, inp, dtype=torch.bfloat16, taps=None):
    f = (lambda t: t.to(dtype) if t.is_floating_point() else t)
    sdd = sd if dtype == torch.bfloat16 else {k: v.float() for k, v in sd.items()}
    tap = None if taps is None else (lambda n, t: taps.__setitem__(n, t.detach().clone()))
    with torch.inference_mode():
        return O.infer_action(sdd, cfg, inp["input_ids"], f(inp["pixel_values"]).clone(),
                              f(inp["image_text_proprio_mask"]), f(inp["action_mask"]), inp["vlm_position_ids"],
                              inp["proprio_position_ids"], inp["action_position_ids"], f(inp["proprios"]),
                              noise=inp["noise"], tap=tap, rope_dtype=torch.bfloat16)


def test_full_model_actions_and_layers(full):
    """Per-layer parity is judged the way two bf16 runs can be judged: both runs (ours, and the
    reference's bf16 op sequence on this GPU) are compared with the fp32 run of the same model, and
    our error must stay within a small factor of the reference's own.  (A direct bf16-vs-bf16 bound in
    ulps is meaningless on the residual streams: elements near zero are sums of O(10) terms.)"""
    cfg, model, sd = full
    inp = synth.synthetic_inputs(cfg, 1, dtype=torch.bfloat16, device=DEV)
    taps, taps32 = {}, {}
    ref = _oracle(sd, cfg, inp, taps=taps)
    ref32 = _oracle(sd, cfg, inp, dtype=torch.float32, taps=taps32)
    model.set_engine_options(debug_taps=True)
    with torch.inference_mode():
        got = model(**synth.call_args(inp), noise=inp["noise"])
    model._engine.check()
    alias = {"merged_embeds": "prefill.embeds.vlm", "flow0.action_embeds": "flow0.embeds.action"}
    names = ["siglip.embeddings"] + [f"siglip.layer{l}" for l in range(27)] + ["siglip.post_layernorm", "projector",
             "merged_embeds"] + [f"prefill.L{l}.{m}" for l in range(17) for m in ("vlm", "proprio")] + \
            ["flow0.action_embeds"] + [f"flow0.L{l}.action" for l in range(18)] + ["flow0.velocity"]
    print("\nper-layer activation error, full-size Bridge: ours vs bf16 reference | each vs the fp32 run (same GPU)")
    worst_ratio = 0.0
    for n in names:
        key = alias.get(n, n)
        r = taps[key].float().flatten()
        hi = taps32[key].float().flatten()
        g = model.debug_tap(n).float().flatten()
        d = (g - r).abs()
        e_ours, e_ref = (g - hi).abs(), (r - hi).abs()
        rms = hi.pow(2).mean().sqrt().item()
        ratio = e_ours.mean().item() / max(e_ref.mean().item(), 1e-4 * rms)
        worst_ratio = max(worst_ratio, ratio)
        print(f"  {n:24s} vs bf16-ref max={d.max().item():.3e} mean={d.mean().item():.3e} mismatch={(d > 0).float().mean().item():.3f}"
              f" | vs fp32: ours max={e_ours.max().item():.3e} mean={e_ours.mean().item():.3e}, bf16-ref max={e_ref.max().item():.3e}"
              f" mean={e_ref.mean().item():.3e} (rms {rms:.3e})")
        assert e_ours.mean().item() <= 1.5 * e_ref.mean().item() + 1e-4 * rms, n
        assert e_ours.max().item() <= 2.0 * e_ref.max().item() + 1e-3 * rms, n
    model.set_engine_options(debug_taps=False)
    err = (got.float() - ref.float()).abs().max().item()
    clamped = (got.float().clamp(-1, 1) - ref.float().clamp(-1, 1)).abs().max().item()
    e_ours, e_ref = (got.float() - ref32).abs().max().item(), (ref.float() - ref32).abs().max().item()
    print(f"actions: ours vs bf16-ref un-clamped {err:.3e}, clamped {clamped:.3e}; vs fp32: ours {e_ours:.3e}, "
          f"bf16-ref {e_ref:.3e}; range [{ref.min().item():.2f}, {ref.max().item():.2f}]; worst per-layer mean-error ratio "
          f"ours/bf16-ref {worst_ratio:.2f}")
    assert torch.isfinite(got.float()).all()
    assert clamped <= 1e-2                      # north_star tolerance
    assert err <= 3.2e-2                        # 2 bf16 ulp at |a| in [2, 4)
    assert e_ours <= e_ref + 1.6e-2             # no worse than the reference's own bf16 error (+1 ulp)


def test_full_model_sharding_invariance_and_kv_layout(full):
    """Episodes are independent: at a given per-GPU batch size an episode's actions depend on its own
    inputs only — not on its slot in the batch nor on which other episodes share the launch — which
    is what makes the 1/2/4/8-GPU episode sharding exact (every rank runs the same batch size; the
    split-K factor, hence the fp32 summation grouping, is a function of the batch size only).
    KV slot i <-> position id i+1."""
    cfg, model, sd = full
    inp = synth.synthetic_inputs(cfg, 4, dtype=torch.bfloat16, vary_text=True, device=DEV)
    other = synth.synthetic_inputs(cfg, 4, seed=77, dtype=torch.bfloat16, vary_text=True, device=DEV)
    perm = torch.tensor([2, 0, 3, 1], device=DEV)

    def pick(d, idx):
        return {k: (v[idx] if k in bdist.BATCH_KEYS else v) for k, v in d.items()}

    def mix(a, b):      # episodes 0,1 of `a` with episodes 2,3 of `b`
        return {k: (torch.cat([a[k][:2], b[k][2:]]) if k in bdist.BATCH_KEYS else a[k]) for k in a}

    with torch.inference_mode():
        whole = model(**synth.call_args(inp), noise=inp["noise"]).clone()
        permuted = model(**synth.call_args(pick(inp, perm)), noise=inp["noise"][perm]).clone()
        mixed_in = mix(inp, other)
        mixed = model(**synth.call_args(mixed_in), noise=mixed_in["noise"]).clone()
        # two "ranks" of a 2-GPU run, each with its own 2 episodes, vs the same 2-episode batches alone
        parts = [model(**synth.call_args(bdist.shard_inputs(inp, 2, r)),
                       noise=bdist.shard_inputs(inp, 2, r)["noise"]).clone() for r in range(2)]
        again = [model(**synth.call_args(bdist.shard_inputs(inp, 2, r)),
                       noise=bdist.shard_inputs(inp, 2, r)["noise"]).clone() for r in range(2)]
        last = bdist.shard_inputs(inp, 4, 3)
        single = model(**synth.call_args(last), noise=last["noise"]).clone()
    model._engine.check()
    assert torch.equal(whole[perm], permuted)                 # slot in the batch does not matter
    assert torch.equal(whole[:2], mixed[:2])                  # companions do not matter
    assert all(torch.equal(a, b) for a, b in zip(parts, again))   # run-to-run deterministic
    assert (torch.cat(parts).float() - whole.float()).abs().max().item() <= 3.2e-2   # other batch size: ~1 ulp
    assert (single.float() - whole[3:].float()).abs().max().item() <= 3.2e-2
    # KV cache of the last call (episode 3 alone): pad slots of the vlm block hold the pad rows'
    # keys, proprio sits at slot 276, nothing is shifted by the shorter text
    L = cfg.joint.config.num_hidden_layers
    kc = model.debug_tap("k_cache").view(L, model._engine.max_batch, 281, 256)
    ref_taps = {}
    loc = bdist.shard_inputs(inp, 4, 3)
    with torch.inference_mode():
        _, caches = O.infer_action(sd, cfg, **synth.call_args(loc), noise=loc["noise"], return_caches=True)
    for l in (0, L // 2, L - 1):
        k_ref = torch.cat([caches["vlm"].key_cache[l], caches["proprio"].key_cache[l]], dim=2)[0, 0]
        e = (kc[l, 0, :277].float() - k_ref.float()).abs()
        print(f"  k_cache L{l}: max_abs={e.max().item():.3e} (ref rms {k_ref.float().pow(2).mean().sqrt():.3f})")
        assert e.max().item() <= 0.13


def _replicate(inp, times):
    return {k: (v.repeat_interleave(times, 0) if k in bdist.BATCH_KEYS else v) for k, v in inp.items()}


def test_full_model_ten_step_vs_single_step(full):
    """BASELINE.json configs[2] on the Bridge weights: 10-step flow vs single step from the same injected noise; ours
    must track the reference op sequence in both schedules.  Tolerance: north_star's 1e-2 on the clamped actions, or —
    where the reference's OWN bf16 run moves by more than that when nothing but its summation order changes (the same
    episode evaluated inside a batch of 8 copies: cuBLAS picks other kernels) — that measured reproducibility floor."""
    cfg, model, sd = full
    inp = synth.synthetic_inputs(cfg, 1, dtype=torch.bfloat16, device=DEV)
    out = {}
    for steps in (10, 1):
        cfg.num_inference_steps = steps
        model.num_inference_steps = steps        # new time table + schedule, weights stay uploaded
        with torch.inference_mode():
            got = model(**synth.call_args(inp), noise=inp["noise"]).float()
        ref = _oracle(sd, cfg, inp).float()
        ref_b8 = _oracle(sd, cfg, _replicate(inp, 8)).float()[:1]
        out[steps] = (got, ref)
        clamped = (got.clamp(-1, 1) - ref.clamp(-1, 1)).abs().max().item()
        floor = (ref_b8.clamp(-1, 1) - ref.clamp(-1, 1)).abs().max().item()
        print(f"steps={steps}: ours vs bf16-ref max_abs {((got - ref).abs().max().item()):.3e}, clamped {clamped:.3e}; "
              f"bf16-ref vs itself inside a batch of 8: clamped {floor:.3e}")
        # bf16 actions in [0.5, 1) are 2^-8 = 3.9e-3 apart: allow the reference's own floor plus one such step
        assert clamped <= max(1e-2, floor + 2 ** -8)
    d_ours = (out[10][0] - out[1][0]).abs().max().item()
    d_ref = (out[10][1] - out[1][1]).abs().max().item()
    print(f"10-step vs 1-step action difference: ours {d_ours:.3e}, reference {d_ref:.3e}")
    assert abs(d_ours - d_ref) <= 5e-2


def test_full_model_stress_weights_reported(full):
    """SURVEY.md 8(c)-5 / BASELINE.md section 3 at full depth: q/k weights x8 (attention logits x64: soft-clamp, masks
    and RoPE matter).  Two bf16 runs with different summation orders diverge on this weight set (the
    reference's own bf16 run is 0.42 away from its fp32 run), so the numbers are REPORTED and the assertion is the
    fp32 tie-break: ours is no further from the fp32 run than the reference's bf16 op sequence is (x1.5)."""
    cfg, model, _ = full
    sd = synth.synthetic_state_dict(cfg, 0, torch.bfloat16, stress=True)
    stress_model = PiZeroInference.from_state_dict(cfg, sd, device=DEV)
    sd_gpu = {k: v.to(DEV) for k, v in sd.items()}
    inp = synth.synthetic_inputs(cfg, 1, dtype=torch.bfloat16, device=DEV)
    with torch.inference_mode():
        got = stress_model(**synth.call_args(inp), noise=inp["noise"]).float()
    stress_model._engine.check()
    ref = _oracle(sd_gpu, cfg, inp).float()
    ref32 = _oracle(sd_gpu, cfg, inp, dtype=torch.float32).float()
    stress_model.release_engine()
    clamped = (got.clamp(-1, 1) - ref.clamp(-1, 1)).abs().max().item()
    e_ours, e_ref = (got - ref32).abs().max().item(), (ref - ref32).abs().max().item()
    print(f"stress weights, full size: ours vs bf16-ref clamped {clamped:.3e} un-clamped {(got - ref).abs().max().item():.3e}; "
          f"vs fp32: ours {e_ours:.3e}, bf16-ref {e_ref:.3e}")
    assert torch.isfinite(got).all()
    assert e_ours <= 1.5 * e_ref + 1.6e-2


def test_full_model_64_episodes_match_oracle(full):
    """BASELINE.json configs[3] at FULL depth: 64 episodes with per-episode instruction lengths in one launch (the
    batched kernels: persistent CTA-pair GEMMs, tcgen05 attention, streaming consumers) against the oracle's bf16 run
    on the same GPU, 8 episodes at a time.
    Over 64 x 28 action values the reference's bf16 path does not reproduce ITSELF to 1e-2 when only its summation
    order changes (measured here: the same episodes at batch 8 and at batch 1; tools/noise_floor.py,
    profiles/r02_bf16_noise_floor.txt: 1.4e-2 clamped over 16 episodes), so the bound is north_star's 1e-2 or that
    measured floor, and the fp32 run breaks the tie: our mean error against it must not exceed the reference's."""
    cfg, model, sd = full
    n = 64
    inp = synth.synthetic_inputs(cfg, n, seed=4242, dtype=torch.bfloat16, vary_text=True, device=DEV)
    model.set_engine_options(reserve_batch=n)
    with torch.inference_mode():
        got = model(**synth.call_args(inp), noise=inp["noise"]).float().clone()
    model._engine.check()

    def sub(lo, hi):
        return {k: (v[lo:hi] if k in bdist.BATCH_KEYS else v) for k, v in inp.items()}
    ref = torch.cat([_oracle(sd, cfg, sub(lo, lo + 8)).float() for lo in range(0, n, 8)])
    ref_one = torch.cat([_oracle(sd, cfg, sub(i, i + 1)).float() for i in range(n)])
    ref32 = torch.cat([_oracle(sd, cfg, sub(lo, lo + 8), dtype=torch.float32).float() for lo in range(0, 16, 8)])
    c = lambda a, b: (a.clamp(-1, 1) - b.clamp(-1, 1)).abs()
    clamped = c(got, ref).max().item()
    floor = c(ref, ref_one).max().item()
    m_ours, m_ref = c(got[:16], ref32).mean().item(), c(ref[:16], ref32).mean().item()
    print(f"64 episodes, full depth: ours vs bf16-ref(bs=8) clamped max_abs {clamped:.3e} (un-clamped {(got - ref).abs().max().item():.3e}); "
          f"bf16-ref(bs=8) vs bf16-ref(bs=1): {floor:.3e}; mean abs error vs fp32 (16 episodes): ours {m_ours:.3e}, bf16-ref {m_ref:.3e}")
    assert torch.isfinite(got).all()
    assert clamped <= max(1e-2, 1.25 * floor)
    assert c(got, ref).mean().item() <= 3e-3
    assert m_ours <= 1.1 * m_ref + 1e-4
