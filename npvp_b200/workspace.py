"""Named scratch buffers, cached by (name, shape, dtype, device).

Repeated forwards with the same shapes reuse the same device memory, which keeps pointers stable
(a prerequisite for CUDA-graph replay) and avoids allocator traffic on the hot path.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


class Workspace:
    def __init__(self, device: torch.device):
        self.device = device
        self._bufs: Dict[Tuple, torch.Tensor] = {}

    def get(self, name: str, shape, dtype) -> torch.Tensor:
        key = (name, tuple(int(s) for s in shape), dtype)
        buf = self._bufs.get(key)
        if buf is None:
            # Buffers of other shapes under the same name stay alive: a captured CUDA graph replays with the raw pointers of
            # the shapes it was captured with (a rollout alternates between full and short blocks), so freeing them here
            # would hand their memory to someone else under a live graph.  ``trim()`` releases everything explicitly.
            buf = torch.empty(key[1], dtype=dtype, device=self.device)
            self._bufs[key] = buf
        return buf

    def trim(self):
        """Release every cached buffer (only when no captured graph that used them will be replayed again)."""
        self._bufs.clear()

    def bf16(self, name, *shape):
        return self.get(name, shape, torch.bfloat16)

    def h16(self, name, dtype, *shape):
        return self.get(name, shape, dtype)

    def f32(self, name, *shape):
        return self.get(name, shape, torch.float32)

    def bytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self._bufs.values())
