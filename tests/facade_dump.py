"""Parser of the per-frame dumps written by cadr_b200/host/tests/facade_scene_test.cpp."""
from __future__ import annotations

import struct

import numpy as np


def parse(path: str) -> list[dict]:
    data = open(path, "rb").read()
    pos, frames = 0, []

    def take(n):
        nonlocal pos
        b = data[pos:pos + n]
        assert len(b) == n, "truncated dump"
        pos += n
        return b

    def u32():
        return struct.unpack("<I", take(4))[0]

    def u64():
        return struct.unpack("<Q", take(8))[0]

    while pos < len(data):
        assert take(8) == b"CADRF002"
        f = dict(frame=u32(), has_device=bool(u32()))
        segs = []
        for _ in range(u32()):
            base, size = u64(), u64()
            segs.append((base, np.frombuffer(take(size), dtype=np.uint8).copy()))
        f["segments"] = segs
        f["root"], f["level"], n = u64(), u32(), u32()
        f["n"] = n
        f["highest_handle"], f["drawable_buffer"] = u64(), u64()
        f["drawables"] = np.frombuffer(take(n * 48), dtype=np.uint64).reshape(n, 6).copy()
        f["cull"] = np.frombuffer(take(n * 48), dtype=np.uint32).reshape(n, 12).copy()
        fr = np.frombuffer(take(108), dtype=np.float32)
        f["planes"], f["eye"] = fr[:24].reshape(6, 4).copy(), fr[24:27].copy()
        f["ranges"] = [dict(first=u32(), count=u32(), ptr_off=u64(), ind_off=u64()) for _ in range(u32())]
        exp = np.frombuffer(take(n * 48), dtype=np.uint8).reshape(n, 48)
        f["expected_indirect"] = exp[:, :16].copy().view(np.uint32).reshape(n, 4)
        f["expected_pointers"] = exp[:, 16:].copy().view(np.uint64).reshape(n, 4)
        R = u32()
        f["regions"] = np.frombuffer(take(R * 16), dtype=np.uint32).reshape(R, 4).copy()
        if f["has_device"]:
            f["gpu_indirect"] = np.frombuffer(take(n * 16), dtype=np.uint32).reshape(n, 4).copy()
            f["gpu_pointers"] = np.frombuffer(take(n * 32), dtype=np.uint64).reshape(n, 4).copy()
            cmd_cap, inst_cap = u64(), u64()
            raw = np.frombuffer(take(64 + 8 * R), dtype=np.uint8).copy()
            hdr, packed = raw[:64].view(np.uint32), raw[64:].view(np.uint64)
            g = dict(status=int(hdr[0]), near_band=int(hdr[1]), regions=f["regions"],
                     cmd_count=(packed & np.uint64(0xFFFFFFFF)).astype(np.int64), inst_count=(packed >> np.uint64(32)).astype(np.int64))
            g["cmd"] = np.frombuffer(take(cmd_cap * 20), dtype=np.uint32).reshape(-1, 5).copy()
            g["ptr"] = np.frombuffer(take(cmd_cap * 32), dtype=np.uint64).reshape(-1, 4).copy()
            g["tag"] = np.frombuffer(take(cmd_cap * 8), dtype=np.uint32).reshape(-1, 2).copy()
            g["inst"] = np.frombuffer(take(inst_cap * 4), dtype=np.uint32).copy()
            f["gpu_cull"] = g
        frames.append(f)
    return frames
