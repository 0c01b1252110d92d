"""CUDA side of the streaming slot attention (contract: devias_b200/slot_attention.py).

INTERIM (round 1, first slice): `slot_stream` / `token_stats` evaluate the folded form with torch ops on the
GPU while the hand-written streaming kernel (csrc/slot_attn.cu) is brought up; the contract is fixed."""
from . import slot_attention as SA

token_stats = SA.token_stats
slot_stream = SA.slot_stream_torch
