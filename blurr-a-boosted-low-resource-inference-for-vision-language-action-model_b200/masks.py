"""Block-attention mask and position ids of the Pi-0 sequence layout
`[image/text (276, right-padded) | proprio (1) | action (4)]`.

Restates `PiZero.build_causal_mask_and_position_ids` / `split_full_mask_into_submasks`
(reference `src/model/vla/pizero.py:328-393`) without the per-sample Python loop (:353-357);
values are bit-identical (tests/test_masks.py checks this against the oracle and the reference).
"""

from __future__ import annotations

from typing import Tuple

import torch


def build_causal_mask_and_position_ids(attention_mask: torch.Tensor, dtype: torch.dtype,
                                       max_image_text_tokens: int, num_proprio_tokens: int,
                                       num_action_tokens: int
                                       ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """
             img/text (cnt valid) | pad | proprio | action
    img/text        0             | min |  min    |  min      (rows >= cnt: all min)
    proprio         0             | min |   0     |  min
    action          0             | min |   0     |   0
    """
    bsz = attention_mask.size(0)
    device = attention_mask.device
    n_it, n_p, n_a = max_image_text_tokens, num_proprio_tokens, num_action_tokens
    total = n_it + n_p + n_a
    cnt = torch.sum(attention_mask, dim=1).view(bsz, 1, 1)
    r = torch.arange(total, device=device).view(1, total, 1)
    c = torch.arange(total, device=device).view(1, 1, total)
    allowed = ((r < cnt) | (r >= n_it)) & (c < cnt)           # pizero.py:354-357
    allowed = allowed | ((r >= n_it) & (r < n_it + n_p) & (c >= n_it) & (c < n_it + n_p))   # :358-360
    allowed = allowed | ((r >= n_it + n_p) & (c >= n_it))     # :361-363
    causal_mask = torch.full((bsz, total, total), torch.finfo(dtype).min, dtype=dtype, device=device)
    causal_mask.masked_fill_(allowed, 0)
    causal_mask = causal_mask.unsqueeze(1)                    # :367
    vlm_position_ids = torch.arange(1, n_it + 1).repeat(bsz, 1)                   # :370-372
    proprio_position_ids = torch.arange(1, n_p + 1).repeat(bsz, 1)                # :373-375
    action_position_ids = torch.arange(n_p + 1, n_p + n_a + 1).repeat(bsz, 1)     # :376-379
    return causal_mask, vlm_position_ids, proprio_position_ids, action_position_ids


def split_full_mask_into_submasks(causal_mask: torch.Tensor, max_image_text_tokens: int,
                                  num_proprio_tokens: int, num_action_tokens: int):
    """pizero.py:383-393: views `[..., :277, :277]` and `[..., -4:, :]`."""
    n = max_image_text_tokens + num_proprio_tokens
    return causal_mask[..., :n, :n], causal_mask[..., -num_action_tokens:, :]
