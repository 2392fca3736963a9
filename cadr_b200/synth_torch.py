"""On-device synthesis of the big configs' matrix lists (torch is plumbing here: device memory + RNG math).

The same counter-based PRNG as cadr_b200/synth.py (splitmix64 of the instance index), so a device-generated
scene has the same distribution — and the same uniform variates bit for bit — as the host recipe; only the
transcendental functions (sin/cos/log) may differ in the last ulp between numpy and CUDA.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .synth import ML_HEADER, Scene


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def u01(seed: int, stream: int, idx: torch.Tensor) -> torch.Tensor:
    key = _s64(seed * 0x100000001B3 + stream * 0x9E3779B97F4A7C15)
    h = splitmix64(idx * _s64(0xD1342543DE82EF95) + key)
    return _lsr(h, 40).to(torch.float32) * (1.0 / (1 << 24))


def gauss(seed: int, stream: int, idx: torch.Tensor) -> torch.Tensor:
    u1 = torch.clamp_min(u01(seed, stream, idx), 1e-7)
    u2 = u01(seed, stream + 1, idx)
    return torch.sqrt(-2 * torch.log(u1)) * torch.cos((2 * math.pi) * u2)


def trs(pos: torch.Tensor, quat: torch.Tensor | None, scale: torch.Tensor) -> torch.Tensor:
    m = pos.shape[0]
    out = torch.zeros((m, 16), dtype=torch.float32, device=pos.device)
    if quat is None:
        out[:, 0] = scale; out[:, 5] = scale; out[:, 10] = scale
    else:
        x, y, z, w = quat.unbind(1)
        s = scale
        out[:, 0] = (1 - 2 * (y * y + z * z)) * s; out[:, 1] = (2 * (x * y + z * w)) * s; out[:, 2] = (2 * (x * z - y * w)) * s
        out[:, 4] = (2 * (x * y - z * w)) * s; out[:, 5] = (1 - 2 * (x * x + z * z)) * s; out[:, 6] = (2 * (y * z + x * w)) * s
        out[:, 8] = (2 * (x * z + y * w)) * s; out[:, 9] = (2 * (y * z - x * w)) * s; out[:, 10] = (1 - 2 * (x * x + y * y)) * s
    out[:, 12:15] = pos
    out[:, 15] = 1
    return out


def c3_matrices(seed: int, lists: torch.Tensor, instances: int, cube: float, sigma: float) -> torch.Tensor:
    k = lists.repeat_interleave(instances)
    j = torch.arange(instances, dtype=torch.int64, device=lists.device).repeat(lists.numel())
    g = k * instances + j
    centre = torch.stack([(u01(seed, s, k) - 0.5) * cube for s in (0, 1, 2)], dim=1)
    pos = centre + sigma * torch.stack([gauss(seed, 10 + 2 * s, g) for s in (0, 1, 2)], dim=1)
    u1, u2, u3 = u01(seed, 20, g), u01(seed, 21, g), u01(seed, 22, g)
    a, b = torch.sqrt(1 - u1), torch.sqrt(u1)
    t2, t3 = (2 * math.pi) * u2, (2 * math.pi) * u3
    q = torch.stack([a * torch.sin(t2), a * torch.cos(t2), b * torch.sin(t3), b * torch.cos(t3)], dim=1)
    scale = 0.5 + 1.5 * u01(seed, 30, g)
    return trs(pos, q, scale)


def c2_matrices(seed: int, idx: torch.Tensor, cube: float) -> torch.Tensor:
    pos = torch.stack([(u01(seed, s, idx) - 0.5) * cube for s in (0, 1, 2)], dim=1)
    scale = 0.5 + 1.5 * u01(seed, 3, idx)
    return trs(pos, None, scale)


class TorchArena:
    """Device allocator backed by torch tensors (injected into cadr_b200.frame.DeviceScene)."""

    def __init__(self, device: torch.device):
        self.device = device
        self.tensors: dict[int, torch.Tensor] = {}

    def alloc(self, nbytes: int) -> int:
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.tensors[t.data_ptr()] = t
        return t.data_ptr()

    def free(self, addr: int) -> None:
        self.tensors.pop(addr, None)

    def tensor(self, addr: int) -> torch.Tensor:
        return self.tensors[addr]


def fill_matrix_lists(scene: Scene, arena: torch.Tensor, slab_lists: int = 2000) -> None:
    """Write every MatrixList block (header + matrices) of a uniform scene (cfg 2 / cfg 3) into `arena`."""
    kind = scene.gen.get("kind")
    L = len(scene.ml_off)
    count = int(scene.ml_count[0])
    assert (scene.ml_count == count).all(), "device synthesis needs lists of equal size"
    stride = ML_HEADER + 64 * count
    start = int(scene.ml_off[0])
    assert L == 1 or int(scene.ml_off[1]) - start == stride
    blocks = arena[start:start + L * stride].view(L, stride)
    hdr = torch.zeros(ML_HEADER, dtype=torch.uint8)
    hdr[:8] = torch.from_numpy(np.array([count, count], dtype=np.uint32).view(np.uint8))
    blocks[:, :ML_HEADER] = hdr.to(arena.device)
    if kind == "c2":
        slab_lists = max(slab_lists, 1 << 20)
    # a shard of a bigger scene (synth.config3_shard): local list k holds the matrices of GLOBAL list list_ids[k]
    ids = scene.gen.get("list_ids")
    ids = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).to(arena.device) if ids is not None else None
    for a in range(0, L, slab_lists):
        b = min(L, a + slab_lists)
        lists = torch.arange(a, b, dtype=torch.int64, device=arena.device) if ids is None else ids[a:b]
        if kind == "c3":
            m = c3_matrices(scene.seed, lists, count, scene.gen["cube"], scene.gen["sigma"])
        elif kind == "c2":
            m = c2_matrices(scene.seed, lists, scene.gen["cube"])
        else:
            raise ValueError(f"no device recipe for scene kind {kind!r}")
        blocks[a:b, ML_HEADER:] = m.view(torch.uint8).view(b - a, count * 64)
