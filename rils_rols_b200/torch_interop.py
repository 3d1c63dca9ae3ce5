"""torch plumbing: wrap raw device pointers handed out by the C ABI as tensors (no copies)."""
from __future__ import annotations

import torch


class _DevArray:
    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


def tensor_from_ptr(ptr: int, count: int, device: torch.device) -> torch.Tensor:
    return torch.as_tensor(_DevArray(ptr, count), device=device)
