"""Per-frame driver over the C ABI: what CadR::Renderer does between beginFrame and the draw calls.

  DeviceScene.record_drawable_processing  == Renderer::recordDrawableProcessing (Renderer.cpp:598-720):
        DMA the flattened DrawableGpuData list, run the processDrawables kernel
  DeviceScene.cull                        == the north-star extension (culling + LOD + compaction)

Buffer sizing follows Renderer::prepareSceneRendering (Renderer.cpp:446-595): drawable / indirect / pointers
buffers hold floor(n*1.2) records, at least 128.  All device memory comes from the context
(cadr_b200_arena_alloc) unless an allocator is injected (bench.py injects a torch-backed one so that it can
synthesise 6.4 GB of matrices on the device).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .synth import Scene

_P = C.c_void_p


def buffer_capacity(n: int) -> int:
    """Renderer.cpp:463-466: 20 % head-room, at least 128 records."""
    return max(int(np.float32(n) * np.float32(1.2)), 128)


class DeviceScene:
    def __init__(self, ctx: _capi.Context, scene: Scene, *, alloc=None, free=None, upload: bool = True,
                 stream: int = 0):
        self.ctx, self.scene, self.stream = ctx, scene, stream
        self._alloc = alloc or ctx.arena_alloc
        self._free = free or ctx.arena_free
        self._owned = []
        n, S = scene.n, scene.num_state_sets
        cap = buffer_capacity(n)
        self.capacity = cap
        self.arena = self._new(scene.arena_bytes)
        self.drawable_list = self._new(cap * 48)
        self.indirect = self._new(cap * 16)
        self.pointers = self._new(cap * 32)
        self.cull_data = self._new(max(n, 1) * 48)
        self.regions = self._new(S * 16)
        self.cmd_cap, self.inst_cap = max(scene.cmd_capacity, 1), max(scene.inst_capacity, 1)
        self.cmd_out = self._new(self.cmd_cap * 20)
        self.ptr_out = self._new(self.cmd_cap * 32)
        self.tag_out = self._new(self.cmd_cap * 8)
        self.inst_out = self._new(self.inst_cap * 4)
        self.counters_bytes = ctx.cull_counters_bytes(S)
        self.counters = self._new(self.counters_bytes)
        self.chunk_cap = scene.chunk_capacity
        self.chunk_ws = self._new(max(self.chunk_cap, 1) * 128)   # CADR_CULL_WORK_ITEM_BYTES
        self.root = self.arena + scene.root_off
        self.bounds = 0               # optional cadr_drawable_bound[n], see compute_bounds()
        # host staging for the drawable list (the reference keeps it in a mapped HOST_CACHED buffer,
        # Renderer.cpp:513-535) — pinned here so the per-frame copy is a true DMA
        self.host_list_ptr = ctx.host_alloc(cap * 48) if ctx.device >= 0 else 0
        if upload:
            self.upload_static()

    def _new(self, nbytes: int) -> int:
        a = self._alloc(max(int(nbytes), 16))
        self._owned.append(a)
        return a

    def close(self) -> None:
        for a in self._owned:
            self._free(a)
        self._owned = []
        if self.host_list_ptr:
            self.ctx.host_free(self.host_list_ptr)
            self.host_list_ptr = 0

    # -- static data -------------------------------------------------------------------------------
    def upload_static(self, with_matrices: bool = True) -> None:
        sc, ctx, s = self.scene, self.ctx, self.stream
        if with_matrices and sc.matrices is not None:
            img = sc.image(self.arena)
            ctx.memcpy_h2d(self.arena, img, stream=s)
        else:
            meta = sc.metadata_extent()
            img = np.zeros(meta, dtype=np.uint8)
            sc.write_metadata(img, self.arena)
            ctx.memcpy_h2d(self.arena, img, stream=s)
        ctx.memcpy_h2d(self.cull_data, np.ascontiguousarray(sc.cull), stream=s)
        ctx.memcpy_h2d(self.regions, np.ascontiguousarray(sc.regions), stream=s)
        if self.host_list_ptr and sc.n:
            C.memmove(self.host_list_ptr, sc.drawables.ctypes.data, sc.n * 48)
        ctx.sync(s)

    # -- per frame ---------------------------------------------------------------------------------
    def record_drawable_processing(self, stream: int | None = None) -> None:
        s = self.stream if stream is None else stream
        self.ctx.record_drawable_processing(self.host_list_ptr, self.root, self.scene.handle_level, self.drawable_list,
                                            self.indirect, self.pointers, self.scene.n, stream=s)

    def process_drawables(self, stream: int | None = None) -> None:
        """Kernel only: the drawable list is already resident (copied by an earlier record_drawable_processing)."""
        s = self.stream if stream is None else stream
        self.ctx.process_drawables(self.root, self.scene.handle_level, self.drawable_list, self.indirect,
                                   self.pointers, self.scene.n, stream=s)

    def cull_params(self, planes: np.ndarray, eye: np.ndarray) -> _capi.CullParams:
        sc = self.scene
        p = _capi.CullParams()
        p.handleTableRoot, p.handleLevel, p.numDrawables = self.root, sc.handle_level, sc.n
        p.drawableList, p.indirectData, p.drawablePointers, p.cullData = self.drawable_list, self.indirect, self.pointers, self.cull_data
        pl = np.ascontiguousarray(planes, dtype=np.float32).reshape(6, 4)
        for k in range(6):
            for c in range(4):
                p.planes[k][c] = float(pl[k, c])
        for c in range(3):
            p.eye[c] = float(eye[c])
        p.numStateSets, p.stateSetRegions = sc.num_state_sets, self.regions
        p.cmdOut, p.ptrOut, p.tagOut, p.instOut, p.counters = self.cmd_out, self.ptr_out, self.tag_out, self.inst_out, self.counters
        p.chunkWorkspace, p.chunkCapacity = self.chunk_ws, self.chunk_cap
        p.drawableBounds = self.bounds
        return p

    def compute_bounds(self, indices: int = 0, count: int | None = None, stream: int | None = None) -> None:
        """Optional pre-test: (re)compute the per-drawable bounds from the current matrices (needs the Tier R outputs of
        the current scene state: run record_drawable_processing / process_drawables first).  Later frames drop long
        lists that lie outside the frustum without reading their matrices."""
        s = self.stream if stream is None else stream
        if not self.bounds:
            self.bounds = self._new(max(self.scene.n, 1) * 32)
        p = self.cull_params(np.zeros((6, 4), np.float32), np.zeros(3, np.float32))
        self.ctx.compute_drawable_bounds(p, self.bounds, self.scene.n if count is None else count, indices, stream=s)

    def cull(self, planes: np.ndarray, eye: np.ndarray, stream: int | None = None) -> None:
        s = self.stream if stream is None else stream
        self.ctx.cull_compact(self.cull_params(planes, eye), stream=s)

    def process_and_cull(self, planes: np.ndarray, eye: np.ndarray, stream: int | None = None) -> None:
        """Tier R + Tier X in one pass over the (already resident) drawable list."""
        s = self.stream if stream is None else stream
        self.ctx.process_and_cull(self.cull_params(planes, eye), stream=s)

    def upload_drawable_list(self, stream: int | None = None) -> None:
        """The per-frame DMA of Renderer::recordDrawableProcessing (Renderer.cpp:635-644) on its own."""
        s = self.stream if stream is None else stream
        self.ctx.memcpy_h2d(self.drawable_list, self.host_list_ptr, self.scene.n * 48, stream=s)

    # -- read-back ---------------------------------------------------------------------------------
    def _read(self, addr: int, nbytes: int, dtype) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        if nbytes:
            self.ctx.memcpy_d2h(out, addr, stream=self.stream)
            self.ctx.sync(self.stream)
        return out.view(dtype)

    def read_tier_r(self) -> tuple[np.ndarray, np.ndarray]:
        n = self.scene.n
        return (self._read(self.indirect, n * 16, np.uint32).reshape(n, 4),
                self._read(self.pointers, n * 32, np.uint64).reshape(n, 4))

    def read_counters(self) -> dict:
        raw = self._read(self.counters, self.counters_bytes, np.uint8)
        hdr = raw[:64].view(np.uint32)
        packed = raw[64:].view(np.uint64)
        # work items queued: long items (hdr[2]) + medium lists (hdr[4], the other end of the same workspace)
        return dict(status=int(hdr[0]), near_band=int(hdr[1]), chunk_count=int(hdr[2]) + int(hdr[4]), medium_count=int(hdr[4]),
                    cmd_count=(packed & np.uint64(0xFFFFFFFF)).astype(np.int64), inst_count=(packed >> np.uint64(32)).astype(np.int64))

    def read_tier_x(self) -> dict:
        out = self.read_counters()
        out["cmd"] = self._read(self.cmd_out, self.cmd_cap * 20, np.uint32).reshape(-1, 5)
        out["ptr"] = self._read(self.ptr_out, self.cmd_cap * 32, np.uint64).reshape(-1, 4)
        out["tag"] = self._read(self.tag_out, self.cmd_cap * 8, np.uint32).reshape(-1, 2)
        out["inst"] = self._read(self.inst_out, self.inst_cap * 4, np.uint32)
        out["regions"] = self.scene.regions
        return out


def canonicalise(result: dict) -> dict:
    """Canonical form of a Tier X result: emission order inside a StateSet is
    nondeterministic and long lists are emitted as several commands (one per work item), so per StateSet
    merge commands by (drawableIndex, lod), sort each merged instance run ascending, sort by key.
    Returns {stateSet: [(drawable, lod, indexCount, firstIndex, vertexOffset, ptr(4), instances ndarray)]}."""
    regions = result["regions"]
    cmd, ptr, tag, inst = result["cmd"], result["ptr"], result["tag"], result["inst"]
    canon = {}
    for s in range(regions.shape[0]):
        base, cnt = int(regions[s, 0]), int(result["cmd_count"][s])
        ibase, icap = int(regions[s, 2]), int(regions[s, 3])
        merged = {}
        for ci in range(base, base + cnt):
            key = (int(tag[ci, 0]), int(tag[ci, 1]))
            k, first = int(cmd[ci, 1]), int(cmd[ci, 4])
            assert ibase <= first and first + k <= ibase + icap, "instance run outside its StateSet region"
            run = inst[first:first + k]
            fixed = (int(cmd[ci, 0]), int(cmd[ci, 2]), int(cmd[ci, 3]), tuple(int(x) for x in ptr[ci]))
            if key in merged:
                assert merged[key][0] == fixed, "commands of one (drawable, lod) disagree"
                merged[key][1].append(run)
            else:
                merged[key] = (fixed, [run])
        lst = []
        for key in sorted(merged):
            fixed, runs = merged[key]
            allinst = np.sort(np.concatenate(runs))
            lst.append((key[0], key[1], fixed[0], fixed[1], fixed[2], fixed[3], allinst))
        canon[s] = lst
        assert sum(len(e[6]) for e in lst) == int(result["inst_count"][s]), "instance count mismatch"
    return canon


def canon_equal(a: dict, b: dict) -> tuple[bool, str]:
    if a.keys() != b.keys():
        return False, "different StateSets"
    for s in a:
        if len(a[s]) != len(b[s]):
            return False, f"StateSet {s}: {len(a[s])} vs {len(b[s])} (drawable, lod) commands"
        for x, y in zip(a[s], b[s]):
            if x[:6] != y[:6]:
                return False, f"StateSet {s}: command {x[:6]} vs {y[:6]}"
            if x[6].shape != y[6].shape or not np.array_equal(x[6], y[6]):
                return False, f"StateSet {s}: instance set of drawable {x[0]} lod {x[1]} differs"
    return True, ""
