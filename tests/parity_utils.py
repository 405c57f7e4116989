"""Shared helpers for the GPU parity tests: run the oracle and the CUDA engine on the same
seeded inputs and compare stage by stage."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from oracle import infinisst_oracle as O


def bf16_weights(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The model lives in bf16 (agents/infinisst.py:150-154,173): both sides get bf16-representable
    weights so the comparison isolates kernel arithmetic."""
    return {k: v.bfloat16().float() for k, v in sd.items()}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.float().flatten(), b.float().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.float() - b.float()).abs().max())


class OracleStream:
    """One stream driven chunk by chunk through the oracle (fp32 or bf16-eager emulation)."""

    def __init__(self, cfg, sd, dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.sd = O.cast_state_dict(sd, dtype)
        self.st = O.StreamState()

    def chunk(self, audio_so_far: List[float], forced: Optional[List[int]] = None):
        taps: dict = {}
        out_ids, rec = O.policy_chunk(self.sd, self.cfg, self.st, audio_so_far, self.dtype, taps, forced)
        return out_ids, rec, taps


def slot_map(cfg, ids: List[int]) -> List[int]:
    return O.speech_slot_map(cfg.llm, ids)


def tie_aware_match(scores: torch.Tensor, token: int, eps: float) -> bool:
    """Accept `token` if the oracle's processed score for it is within eps of the oracle's max
    (SURVEY §7 hard part 2: near-ties under random weights)."""
    return bool(scores[token] >= scores.max() - eps)
