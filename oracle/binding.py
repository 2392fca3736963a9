"""ctypes binding of oracle/libcadr_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (cadr_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcadr_oracle.so")


class Segment(C.Structure):
    _fields_ = [("base", C.c_uint64), ("bytes", C.c_uint64), ("host", C.c_void_p)]


class CullResult(C.Structure):
    _fields_ = [("numCommands", C.c_uint64), ("numInstances", C.c_uint64), ("nearBand", C.c_uint64),
                ("faults", C.c_uint64), ("overflow", C.c_uint32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        l = C.CDLL(LIB_PATH)
        l.oracle_process_drawables.restype = C.c_uint64
        l.oracle_process_drawables.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_uint64, C.c_int, C.c_uint64,
                                               C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        l.oracle_cull_compact.restype = None
        l.oracle_cull_compact.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_uint64, C.c_int, C.c_uint64, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(CullResult)]
        l.oracle_cull_count.restype = C.c_uint64
        l.oracle_cull_count.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
        l.oracle_cull_summary.restype = C.c_uint64
        l.oracle_cull_summary.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
        l.oracle_upload.restype = C.c_uint64
        l.oracle_upload.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        l.oracle_patch_handles.restype = C.c_uint64
        l.oracle_patch_handles.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32]
        l.oracle_consume_check.restype = C.c_uint64
        l.oracle_consume_check.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        l.oracle_consume_check_culled.restype = C.c_uint64
        l.oracle_consume_check_culled.argtypes = [C.POINTER(Segment), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_uint32, C.c_uint32, C.c_void_p]
        l.oracle_max_threads.restype = C.c_int
        l.oracle_transform_spheres.restype = None
        l.oracle_transform_spheres.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib = l
    return _lib


def max_threads() -> int:
    return lib().oracle_max_threads()


def transform_spheres(matrices: np.ndarray, spheres: np.ndarray) -> np.ndarray:
    """World-space bounding spheres {cx, cy, cz, r} of n (column-major mat4, model-space sphere) pairs, exactly as the
    Tier X evaluation computes them (the restatement of BoundingSphere.h:70-87)."""
    m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 16)
    b = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
    assert len(m) == len(b)
    out = np.empty((len(m), 4), np.float32)
    lib().oracle_transform_spheres(m.ctypes.data, b.ctypes.data, len(m), out.ctypes.data)
    return out


class Memory:
    """Device memory model: segments {device base address, host mirror}."""

    def __init__(self, segments: list[tuple[int, np.ndarray]]):
        self.arrays = [np.ascontiguousarray(a).view(np.uint8).reshape(-1) for _, a in segments]
        self.segs = (Segment * len(segments))()
        for i, ((base, _), arr) in enumerate(zip(segments, self.arrays)):
            self.segs[i].base, self.segs[i].bytes, self.segs[i].host = base, arr.nbytes, arr.ctypes.data

    @property
    def n(self) -> int:
        return len(self.arrays)


def process_drawables(mem: Memory, root: int, level: int, drawable_list: int, n: int, threads: int = 1):
    """-> (indirect uint32 [n,4], pointers uint64 [n,4]).  Raises on a memory fault (UB in the reference)."""
    ind = np.zeros((max(n, 1), 4), dtype=np.uint32)
    ptr = np.zeros((max(n, 1), 4), dtype=np.uint64)
    faults = lib().oracle_process_drawables(mem.segs, mem.n, root, level, drawable_list, n, ind.ctypes.data,
                                            ptr.ctypes.data, threads)
    if faults:
        raise RuntimeError(f"oracle: {faults} drawables touched memory outside the arena")
    return ind[:n], ptr[:n]


def cull_compact(mem: Memory, root: int, level: int, drawable_list: int, n: int, indirect: np.ndarray,
                 pointers: np.ndarray, cull: np.ndarray, planes: np.ndarray, eye: np.ndarray, regions: np.ndarray) -> dict:
    S = regions.shape[0]
    cmd_cap = max(int(regions[:, 1].astype(np.int64).sum()), 1)
    inst_cap = max(int(regions[:, 3].astype(np.int64).sum()), 1)
    cmd = np.zeros((cmd_cap, 5), dtype=np.uint32)
    ptr = np.zeros((cmd_cap, 4), dtype=np.uint64)
    tag = np.zeros((cmd_cap, 2), dtype=np.uint32)
    inst = np.zeros(inst_cap, dtype=np.uint32)
    counts = np.zeros(S, dtype=np.uint64)
    res = CullResult()
    ind = np.ascontiguousarray(indirect, dtype=np.uint32)
    pts = np.ascontiguousarray(pointers, dtype=np.uint64)
    cd = np.ascontiguousarray(cull, dtype=np.uint32)
    pl = np.ascontiguousarray(planes, dtype=np.float32).reshape(6, 4)
    ey = np.ascontiguousarray(eye, dtype=np.float32).reshape(-1)[:3].copy()
    rg = np.ascontiguousarray(regions, dtype=np.uint32)
    lib().oracle_cull_compact(mem.segs, mem.n, root, level, drawable_list, n, ind.ctypes.data, pts.ctypes.data,
                              cd.ctypes.data, pl.ctypes.data, ey.ctypes.data, rg.ctypes.data, S,
                              cmd.ctypes.data, ptr.ctypes.data, tag.ctypes.data, inst.ctypes.data,
                              counts.ctypes.data, C.byref(res))
    if res.faults:
        raise RuntimeError(f"oracle: {res.faults} faults in cull_compact")
    return dict(cmd=cmd, ptr=ptr, tag=tag, inst=inst, regions=rg,
                cmd_count=(counts & np.uint64(0xFFFFFFFF)).astype(np.int64),
                inst_count=(counts >> np.uint64(32)).astype(np.int64),
                near_band=int(res.nearBand), status=int(res.overflow),
                num_commands=int(res.numCommands), num_instances=int(res.numInstances))


def cull_count(mem: Memory, first: int, count: int, indirect: np.ndarray, pointers: np.ndarray, cull: np.ndarray,
               planes: np.ndarray, eye: np.ndarray, threads: int) -> tuple[int, int]:
    """Timing variant: -> (survivors, instances visited)."""
    ind = np.ascontiguousarray(indirect, dtype=np.uint32)
    pts = np.ascontiguousarray(pointers, dtype=np.uint64)
    cd = np.ascontiguousarray(cull, dtype=np.uint32)
    pl = np.ascontiguousarray(planes, dtype=np.float32).reshape(6, 4)
    ey = np.ascontiguousarray(eye, dtype=np.float32).reshape(-1)[:3].copy()
    visited = C.c_uint64()
    surv = lib().oracle_cull_count(mem.segs, mem.n, first, count, ind.ctypes.data, pts.ctypes.data, cd.ctypes.data,
                                   pl.ctypes.data, ey.ctypes.data, C.byref(visited), threads)
    return int(surv), int(visited.value)


def host_threads() -> int:
    """Every host core this process may run on (not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1)."""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or max_threads())


def cull_summary(mem: Memory, indirect: np.ndarray, pointers: np.ndarray, cull: np.ndarray, planes: np.ndarray,
                 eye: np.ndarray, threads: int = 0) -> dict:
    """Whole-scene parity: per (drawable, lod) survivor count, index sum and index sum of squares of the drawables
    described by the parallel arrays -> {k [n,3] u32, sum [n,3] u64, sq [n,3] u64, near_band}."""
    ind = np.ascontiguousarray(indirect, dtype=np.uint32)
    pts = np.ascontiguousarray(pointers, dtype=np.uint64)
    cd = np.ascontiguousarray(cull, dtype=np.uint32)
    n = ind.shape[0]
    assert pts.shape[0] == n and cd.shape[0] == n
    pl = np.ascontiguousarray(planes, dtype=np.float32).reshape(6, 4)
    ey = np.ascontiguousarray(eye, dtype=np.float32).reshape(-1)[:3].copy()
    k = np.zeros((max(n, 1), 3), np.uint32)
    sm = np.zeros((max(n, 1), 3), np.uint64)
    sq = np.zeros((max(n, 1), 3), np.uint64)
    nb = C.c_uint64()
    faults = lib().oracle_cull_summary(mem.segs, mem.n, n, ind.ctypes.data, pts.ctypes.data, cd.ctypes.data, pl.ctypes.data,
                                       ey.ctypes.data, k.ctypes.data, sm.ctypes.data, sq.ctypes.data, C.byref(nb),
                                       threads or host_threads())
    if faults:
        raise RuntimeError(f"oracle: {faults} matrix lists outside device memory")
    return dict(k=k[:n], sum=sm[:n], sq=sq[:n], near_band=int(nb.value))


def upload(mem: Memory, regions: np.ndarray, staging: np.ndarray) -> None:
    r = np.ascontiguousarray(regions, dtype=np.uint64).reshape(-1, 3)
    st = np.ascontiguousarray(staging).view(np.uint8).reshape(-1)
    faults = lib().oracle_upload(mem.segs, mem.n, r.ctypes.data, r.shape[0], st.ctypes.data)
    if faults:
        raise RuntimeError(f"oracle: {faults} upload regions outside the arena")


def patch_handles(mem: Memory, root: int, level: int, patches: np.ndarray) -> None:
    p = np.ascontiguousarray(patches, dtype=np.uint64).reshape(-1, 2)
    faults = lib().oracle_patch_handles(mem.segs, mem.n, root, level, p.ctypes.data, p.shape[0])
    if faults:
        raise RuntimeError(f"oracle: {faults} handle patches outside the arena")


def consume_check(mem: Memory, indirect: np.ndarray, pointers: np.ndarray, first: int, count: int) -> tuple[int, int]:
    """Tier R consumer walk -> (digest, fetches)."""
    ind = np.ascontiguousarray(indirect, dtype=np.uint32)
    ptr = np.ascontiguousarray(pointers, dtype=np.uint64)
    out = np.zeros(2, np.uint64)
    faults = lib().oracle_consume_check(mem.segs, mem.n, ind.ctypes.data, ptr.ctypes.data, first, count, out.ctypes.data)
    if faults:
        raise RuntimeError(f"oracle: {faults} draws fetched outside device memory")
    return int(out[0]), int(out[1])


def consume_check_culled(mem: Memory, res: dict, rng: int) -> tuple[int, int]:
    """Tier X consumer walk over draw range `rng` of a cull result -> (digest, fetches)."""
    cmd = np.ascontiguousarray(res["cmd"], dtype=np.uint32)
    ptr = np.ascontiguousarray(res["ptr"], dtype=np.uint64)
    tag = np.ascontiguousarray(res["tag"], dtype=np.uint32)
    inst = np.ascontiguousarray(res["inst"], dtype=np.uint32)
    out = np.zeros(2, np.uint64)
    faults = lib().oracle_consume_check_culled(mem.segs, mem.n, cmd.ctypes.data, ptr.ctypes.data, tag.ctypes.data, inst.ctypes.data,
                                               int(res["regions"][rng, 0]), int(res["cmd_count"][rng]), out.ctypes.data)
    if faults:
        raise RuntimeError(f"oracle: {faults} commands fetched outside device memory")
    return int(out[0]), int(out[1])
