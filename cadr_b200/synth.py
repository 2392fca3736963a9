"""Synthetic scenes for BASELINE.json's configs — input generators for tests and bench.

A Scene is a byte-exact description of what CadR's host side leaves in device memory before
Renderer::recordDrawableProcessing runs:

  * handle tables (2048 x u64 nodes, 1-3 levels)           src/CadR/HandleTable.{h,cpp}
  * per Geometry: vertex block, index block, PrimitiveSet[] src/CadR/Geometry.h:37-39, PrimitiveSet.h:12-15
  * per MatrixList: 64-B header {numMatrices, capacity, 0...} + N x mat4   src/CadR/MatrixList.h:54-59
  * the flattened DrawableGpuData list (48 B each)          src/CadR/Drawable.h:32-43, StateSet.cpp:233-264

Placement follows the reference's packing rule (alignment 64 for sizes >= 64 else 16, successive
allocations packed: src/CadR/CircularAllocationMemory.h:529-537,564-568) inside ONE arena; handle numbering
follows construction order (three DataAllocations per Geometry, then one per MatrixList:
examples/RenderingPerformance/Tests.cpp:484-515,574-578).  Handle tables hold absolute addresses, so the
image is materialised for a given arena base address.

Nothing here computes results; it only lays out inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

TABLE_ENTRIES = 2048
TABLE_BYTES = TABLE_ENTRIES * 8
DRAWABLE_BYTES = 48
CULL_BYTES = 48
ML_HEADER = 64
MAT_BYTES = 64
SMALL_MAX = 32       # mirrors cadr_b200/csrc/cull_compact.cu
CHUNK = 1024


def align_up(x: int, a: int) -> int:
    return (x + a - 1) & ~(a - 1)


class Bump:
    """Packing rule of CircularAllocationMemory::allocPropose (block 1 only; no wrap)."""

    def __init__(self, start: int = 0):
        self.top = start

    def alloc(self, size: int) -> int:
        a = 64 if size >= 64 else 16
        off = align_up(self.top, a)
        self.top = off + size
        return off

    def alloc_array(self, count: int, size: int) -> tuple[int, int]:
        """`count` equal allocations in a row -> (offset of the first, stride)."""
        a = 64 if size >= 64 else 16
        first = align_up(self.top, a)
        stride = align_up(size, a)
        self.top = first + stride * (count - 1) + size if count else self.top
        return first, stride


# ---------------------------------------------------------------------------------------------------
# counter-based PRNG (splitmix64): identical on numpy and torch, any index order
# ---------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def u01(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    """Uniform [0,1) float32 from (seed, stream, index) — 24 random bits."""
    with np.errstate(over="ignore"):
        key = np.uint64((seed * 0x100000001B3 + stream * 0x9E3779B97F4A7C15) & _M64)
        h = splitmix64(idx.astype(np.uint64) * np.uint64(0xD1342543DE82EF95) + key)
    return ((h >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / (1 << 24))


# ---------------------------------------------------------------------------------------------------
# cameras and frustum planes (conventions of the reference: LH, depth 0..1 — glm::perspectiveLH_ZO,
# glm::orthoLH_ZO, glm::lookAtLH; examples/RenderingPerformance/main.cpp:936-946,1416-1421)
# ---------------------------------------------------------------------------------------------------
def look_at_lh(eye, center, up) -> np.ndarray:
    eye, center, up = (np.asarray(v, dtype=np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(up, f)
    s /= np.linalg.norm(s)
    u = np.cross(f, s)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, f
    m[0, 3], m[1, 3], m[2, 3] = -s.dot(eye), -u.dot(eye), -f.dot(eye)
    return m


def perspective_lh_zo(fovy: float, aspect: float, near: float, far: float) -> np.ndarray:
    t = math.tan(fovy / 2)
    m = np.zeros((4, 4))
    m[0, 0] = 1 / (aspect * t)
    m[1, 1] = 1 / t
    m[2, 2] = far / (far - near)
    m[3, 2] = 1
    m[2, 3] = -(far * near) / (far - near)
    return m


def ortho_lh_zo(l, r, b, t, near, far) -> np.ndarray:
    m = np.eye(4)
    m[0, 0], m[1, 1], m[2, 2] = 2 / (r - l), 2 / (t - b), 1 / (far - near)
    m[0, 3], m[1, 3], m[2, 3] = -(r + l) / (r - l), -(t + b) / (t - b), -near / (far - near)
    return m


def frustum_planes(proj: np.ndarray, view: np.ndarray) -> np.ndarray:
    """Six inward-facing unit-normal planes (nx,ny,nz,d), float32, from proj*view (Gribb/Hartmann, depth 0..1)."""
    pv = proj @ view
    rows = [pv[3] + pv[0], pv[3] - pv[0], pv[3] + pv[1], pv[3] - pv[1], pv[2], pv[3] - pv[2]]
    out = np.zeros((6, 4), dtype=np.float32)
    for k, r in enumerate(rows):
        out[k] = (r / np.linalg.norm(r[:3])).astype(np.float32)
    return out


def orbit_camera(frame: int, radius: float, fovy_deg=60.0, aspect=16 / 9, near=0.5, far=1500.0):
    """Appendix D cfg 2/3: eye on a circle around the origin, 1 degree per frame, looking at the origin."""
    a = math.radians(frame % 360)
    eye = np.array([radius * math.sin(a), 0.0, -radius * math.cos(a)])
    view = look_at_lh(eye, (0, 0, 0), (0, 1, 0))
    proj = perspective_lh_zo(math.radians(fovy_deg), aspect, near, far)
    return frustum_planes(proj, view), eye.astype(np.float32)


# ---------------------------------------------------------------------------------------------------
# scene description
# ---------------------------------------------------------------------------------------------------
@dataclass
class Scene:
    name: str
    arena_bytes: int
    handle_level: int
    root_off: int
    num_handles: int
    # handle tables: node offsets and the entries they hold (as arena OFFSETS; 0 == null)
    tables: list = field(default_factory=list)       # [(node_off, idx_array(uint32), target_off_array(uint64))]
    # geometry blocks: [(offset, bytes ndarray)]
    blobs: list = field(default_factory=list)
    # matrix lists
    ml_off: np.ndarray = None        # uint64 [L] offset of each list's header
    ml_count: np.ndarray = None      # uint32 [L]
    matrices: np.ndarray | None = None  # float32 [sum(count), 16] in list order (None: generated on device)
    # per-drawable arrays, flatten order
    drawables: np.ndarray = None     # uint64 [n, 6]: five handles + (psOffset | pad<<32)
    cull: np.ndarray = None          # uint32 [n, 12] (floats bit-cast)
    drawable_ml: np.ndarray = None   # uint32 [n] matrix-list index of each drawable
    drawable_geom: np.ndarray = None # int64 [n] geometry index of each drawable
    geo_off: np.ndarray = None       # uint64 [G, 3] offsets of each geometry's vertex / index / primitive-set block
    dd_off: np.ndarray = None        # uint64 [n] offset of the drawable's data block (0: none)
    regions: np.ndarray = None       # uint32 [S, 4]: cmdBase, cmdCap, instBase, instCap
    seed: int = 0
    gen: dict = field(default_factory=dict)   # generator parameters for on-device matrix synthesis

    @property
    def n(self) -> int:
        return int(self.drawables.shape[0])

    @property
    def num_state_sets(self) -> int:
        return int(self.regions.shape[0])

    @property
    def total_instances(self) -> int:
        return int(self.ml_count[self.drawable_ml].astype(np.int64).sum())

    @property
    def cmd_capacity(self) -> int:
        return int(self.regions[:, 1].astype(np.int64).sum())

    @property
    def inst_capacity(self) -> int:
        return int(self.regions[:, 3].astype(np.int64).sum())

    @property
    def chunk_capacity(self) -> int:
        c = self.ml_count[self.drawable_ml].astype(np.int64)
        big = c[c > SMALL_MAX]
        return int(((big + CHUNK - 1) // CHUNK).sum())

    # -- materialisation -----------------------------------------------------------------------------
    def image(self, base: int, with_matrices: bool = True) -> np.ndarray:
        """The arena bytes for an arena that starts at device address `base`."""
        img = np.zeros(self.arena_bytes, dtype=np.uint8)
        self.write_metadata(img, base)
        if with_matrices:
            if self.matrices is None:
                raise ValueError("scene has no host matrices (device-generated)")
            self.write_matrix_lists(img)
        return img

    def write_metadata(self, img: np.ndarray, base: int) -> None:
        u64 = img.view(np.uint64)
        for node_off, idx, target in self.tables:
            vals = np.where(target != 0, target + np.uint64(base), np.uint64(0))
            u64[node_off // 8 + idx.astype(np.int64)] = vals
        for off, data in self.blobs:
            img[off:off + data.nbytes] = data.view(np.uint8).reshape(-1)

    def write_matrix_lists(self, img: np.ndarray) -> None:
        u32 = img.view(np.uint32)
        u32[self.ml_off.astype(np.int64) // 4] = self.ml_count
        u32[self.ml_off.astype(np.int64) // 4 + 1] = self.ml_count
        f32 = img.view(np.float32)
        start = 0
        # vectorised when all lists have the same size and stride
        for k in range(len(self.ml_off)):
            c = int(self.ml_count[k])
            if c:
                o = (int(self.ml_off[k]) + ML_HEADER) // 4
                f32[o:o + 16 * c] = self.matrices[start:start + c].reshape(-1)
                start += c

    def metadata_extent(self) -> int:
        """Bytes at the start of the arena that hold everything except matrix lists."""
        return int(self.gen.get("metadata_bytes", self.arena_bytes))


def _handle_level(max_handle: int) -> int:
    return 1 if max_handle < TABLE_ENTRIES else (2 if max_handle < TABLE_ENTRIES ** 2 else 3)


def _build_tables(bump: Bump, handles: np.ndarray, targets: np.ndarray, level: int):
    """Sparse handle tables holding `targets[i]` (arena offsets) for `handles[i]`.  Returns (root_off, tables)."""
    # handle 0 ("no drawable data") is looked up like any other handle (processDrawables.comp:111), so its
    # path root[0] -> [mid 0 ->] leaf 0 must exist; slot 0 itself stays zero (HandleTable.cpp:40-46)
    handles = np.concatenate([[0], handles.astype(np.uint64)]).astype(np.uint64)
    targets = np.concatenate([[0], targets.astype(np.uint64)]).astype(np.uint64)
    tables = []
    if level == 1:
        root = bump.alloc(TABLE_BYTES)
        tables.append((root, handles.astype(np.uint32), targets.astype(np.uint64)))
        return root, tables
    leaf_ids, leaf_inv = np.unique(handles >> np.uint64(11), return_inverse=True)
    root = bump.alloc(TABLE_BYTES)
    if level == 2:
        leaf_first, leaf_stride = bump.alloc_array(len(leaf_ids), TABLE_BYTES)
        leaf_off = np.uint64(leaf_first) + np.arange(len(leaf_ids), dtype=np.uint64) * np.uint64(leaf_stride)
        tables.append((root, leaf_ids.astype(np.uint32), leaf_off))
    else:
        mid_ids, mid_inv = np.unique(leaf_ids >> np.uint64(11), return_inverse=True)
        mid_first, mid_stride = bump.alloc_array(len(mid_ids), TABLE_BYTES)
        mid_off = np.uint64(mid_first) + np.arange(len(mid_ids), dtype=np.uint64) * np.uint64(mid_stride)
        leaf_first, leaf_stride = bump.alloc_array(len(leaf_ids), TABLE_BYTES)
        leaf_off = np.uint64(leaf_first) + np.arange(len(leaf_ids), dtype=np.uint64) * np.uint64(leaf_stride)
        tables.append((root, mid_ids.astype(np.uint32), mid_off))
        # one scatter covering all mid nodes: absolute u64 index = mid_off/8 + (leaf_id & 0x7ff)
        tables.append((0, ((mid_off[mid_inv] // np.uint64(8)) + (leaf_ids & np.uint64(0x7FF))).astype(np.uint64),
                       leaf_off))
    # one scatter covering all leaves
    tables.append((0, ((leaf_off[leaf_inv] // np.uint64(8)) + (handles & np.uint64(0x7FF))).astype(np.uint64),
                   targets.astype(np.uint64)))
    return root, tables


def build_scene(name: str, *, geometries: list[dict] | dict, ml_count: np.ndarray, drawable_geom: np.ndarray,
                drawable_ml: np.ndarray, drawable_ps_offset: np.ndarray, state_set: np.ndarray,
                sphere: np.ndarray, lod_count: np.ndarray, lod_ps_offset: np.ndarray, lod_threshold: np.ndarray,
                matrices: np.ndarray | None, drawable_data: np.ndarray | None = None,
                first_handle: int = 1, force_level: int = 0, seed: int = 0, gen: dict | None = None,
                num_state_sets: int = 0) -> Scene:
    """Lay a scene out.

    geometries      either a list of {vertices: bytes ndarray, indices: bytes ndarray, primitive_sets: (P,2) u32},
                    or a dict {count, vertex_bytes, index_bytes, primitive_sets} for `count` identical ones
    drawable_data   optional [n] bool: the drawable owns a 64-byte per-drawable data block (its own handle)
    first_handle    handle numbering starts here (as if first_handle-1 handles had been created before)
    num_state_sets  size of the region table when it is larger than the highest StateSet index used here (a shard of a
                    bigger scene: StateSet indices are global, absent StateSets get empty regions)
    """
    bump = Bump(64)  # offset 0 is never handed out: a zero table entry means "null"
    n = len(drawable_geom)
    L = len(ml_count)
    uniform = isinstance(geometries, dict)
    G = geometries["count"] if uniform else len(geometries)
    num_dd = int(drawable_data.sum()) if drawable_data is not None else 0
    num_handles = 3 * G + L + num_dd
    max_handle = first_handle + num_handles - 1
    level = max(_handle_level(max_handle), force_level)

    # ---- geometry blocks (vertex, index, primitive sets per geometry, in that order) ---------------
    blobs = []
    if uniform:
        vb, ib = int(geometries["vertex_bytes"]), int(geometries["index_bytes"])
        ps = np.ascontiguousarray(geometries["primitive_sets"], dtype=np.uint32)
        # identical geometries: each triple packed one after another
        t0 = bump.top
        v0 = bump.alloc(vb); i0 = bump.alloc(ib); p0 = bump.alloc(ps.nbytes)
        stride = align_up(bump.top - align_up(t0, 64), 64)
        base0 = align_up(t0, 64)
        v_off = np.uint64(v0) + np.arange(G, dtype=np.uint64) * np.uint64(stride)
        i_off = np.uint64(i0) + np.arange(G, dtype=np.uint64) * np.uint64(stride)
        p_off = np.uint64(p0) + np.arange(G, dtype=np.uint64) * np.uint64(stride)
        bump.top = base0 + stride * G
        rng = np.random.default_rng(seed + 17)
        one = np.zeros(stride, dtype=np.uint8)
        # real geometry when given (the consumer-side walk dereferences indices and vertices), else opaque bytes
        vbytes = geometries.get("vertices")
        ibytes = geometries.get("indices")
        one[v0 - base0:v0 - base0 + vb] = rng.integers(0, 256, vb, dtype=np.uint8) if vbytes is None else np.ascontiguousarray(vbytes).view(np.uint8).reshape(-1)
        one[i0 - base0:i0 - base0 + ib] = rng.integers(0, 256, ib, dtype=np.uint8) if ibytes is None else np.ascontiguousarray(ibytes).view(np.uint8).reshape(-1)
        one[p0 - base0:p0 - base0 + ps.nbytes] = ps.view(np.uint8).reshape(-1)
        blobs.append((base0, np.tile(one, G)))
    else:
        v_off = np.zeros(G, dtype=np.uint64); i_off = np.zeros(G, dtype=np.uint64); p_off = np.zeros(G, dtype=np.uint64)
        for g, geo in enumerate(geometries):
            for arr, key in ((v_off, "vertices"), (i_off, "indices"), (p_off, "primitive_sets")):
                data = np.ascontiguousarray(geo[key]).view(np.uint8).reshape(-1)
                arr[g] = bump.alloc(data.nbytes)
                blobs.append((int(arr[g]), data))

    # ---- per-drawable data blocks ------------------------------------------------------------------
    dd_off = np.zeros(0, dtype=np.uint64)
    if num_dd:
        first, stride = bump.alloc_array(num_dd, 64)
        dd_off = np.uint64(first) + np.arange(num_dd, dtype=np.uint64) * np.uint64(stride)
        rng = np.random.default_rng(seed + 23)
        blobs.append((first, rng.integers(0, 256, stride * (num_dd - 1) + 64, dtype=np.uint8)))

    metadata_bytes_before_tables = bump.top

    # ---- matrix lists ------------------------------------------------------------------------------
    ml_count = np.ascontiguousarray(ml_count, dtype=np.uint32)
    sizes = ML_HEADER + MAT_BYTES * ml_count.astype(np.int64)
    # all sizes are multiples of 64 -> each list starts at the 64-aligned end of the previous one
    # (tables are allocated first so that metadata sits at the front of the arena)
    # handles
    h_geo = np.uint64(first_handle) + np.arange(3 * G, dtype=np.uint64)
    h_ml = np.uint64(first_handle + 3 * G) + np.arange(L, dtype=np.uint64)
    h_dd = np.uint64(first_handle + 3 * G + L) + np.arange(num_dd, dtype=np.uint64)

    # reserve tables now; targets need ml offsets, which depend on where the tables end -> two passes
    probe = Bump(bump.top)
    all_handles = np.concatenate([h_geo, h_ml, h_dd])
    _build_tables(probe, all_handles, np.zeros(len(all_handles), dtype=np.uint64), level)
    ml_start = align_up(probe.top, 64)
    ml_off = np.uint64(ml_start) + np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64) if L else np.zeros(0, np.uint64)
    arena_bytes = int(ml_start + sizes.sum())

    geo_targets = np.stack([v_off, i_off, p_off], axis=1).reshape(-1)
    targets = np.concatenate([geo_targets, ml_off, dd_off])
    root_off, tables = _build_tables(bump, all_handles, targets, level)
    assert bump.top == probe.top

    # ---- drawable list -----------------------------------------------------------------------------
    dg = np.asarray(drawable_geom, dtype=np.int64)
    dm = np.asarray(drawable_ml, dtype=np.int64)
    d = np.zeros((n, 6), dtype=np.uint64)
    d[:, 0] = h_geo[3 * dg + 0]
    d[:, 1] = h_geo[3 * dg + 1]
    d[:, 2] = h_ml[dm]
    if num_dd:
        d[np.asarray(drawable_data, dtype=bool), 3] = h_dd
    d[:, 4] = h_geo[3 * dg + 2]
    d[:, 5] = np.asarray(drawable_ps_offset, dtype=np.uint64)

    # ---- culling records ---------------------------------------------------------------------------
    c = np.zeros((n, 12), dtype=np.uint32)
    c[:, 0:4] = np.ascontiguousarray(sphere, dtype=np.float32).view(np.uint32).reshape(n, 4)
    c[:, 4] = lod_count
    c[:, 5:8] = lod_ps_offset
    c[:, 8:10] = np.ascontiguousarray(lod_threshold, dtype=np.float32).view(np.uint32).reshape(n, 2)
    ss = np.asarray(state_set, dtype=np.uint32)
    c[:, 10] = ss

    # ---- per-StateSet output regions, sized for the worst case -------------------------------------
    S = max(int(ss.max()) + 1 if n else 1, int(num_state_sets))
    cnt = ml_count[dm].astype(np.int64)
    cmds = np.where(cnt > SMALL_MAX, 3 * ((cnt + CHUNK - 1) // CHUNK), np.minimum(cnt, 3))
    cmd_cap = np.bincount(ss, weights=cmds, minlength=S).astype(np.int64)
    inst_cap = np.bincount(ss, weights=cnt, minlength=S).astype(np.int64)
    regions = np.zeros((S, 4), dtype=np.uint32)
    regions[:, 1] = cmd_cap
    regions[:, 3] = inst_cap
    regions[:, 0] = np.concatenate([[0], np.cumsum(cmd_cap)[:-1]])
    regions[:, 2] = np.concatenate([[0], np.cumsum(inst_cap)[:-1]])

    dd_per_drawable = np.zeros(n, dtype=np.uint64)
    if num_dd:
        dd_per_drawable[np.asarray(drawable_data, dtype=bool)] = dd_off
    g = dict(gen or {})
    g["metadata_bytes"] = ml_start
    # what slice_scene() needs to lay out a part of this scene again (references, nothing is copied)
    g["build"] = dict(geometries=geometries, ml_count=ml_count, drawable_geom=dg, drawable_ml=dm,
                      drawable_ps_offset=np.asarray(drawable_ps_offset, dtype=np.uint64), state_set=ss,
                      sphere=np.ascontiguousarray(sphere, dtype=np.float32), lod_count=np.asarray(lod_count, dtype=np.uint32),
                      lod_ps_offset=np.asarray(lod_ps_offset, dtype=np.uint32), lod_threshold=np.ascontiguousarray(lod_threshold, dtype=np.float32),
                      drawable_data=drawable_data, first_handle=first_handle, force_level=force_level, num_state_sets=S)
    return Scene(name=name, arena_bytes=arena_bytes, handle_level=level, root_off=root_off, num_handles=num_handles,
                 tables=tables, blobs=blobs, ml_off=ml_off, ml_count=ml_count, matrices=matrices,
                 drawables=d, cull=c, drawable_ml=dm.astype(np.uint32), drawable_geom=dg,
                 geo_off=np.stack([v_off, i_off, p_off], axis=1), dd_off=dd_per_drawable,
                 regions=regions, seed=seed, gen=g)


def slice_scene(scene: Scene, first: int, count: int) -> Scene:
    """One rank's part of ONE scene: drawables [first, first + count) of the flattened list, the geometries
    and matrix lists they use (shared ones once), and a handle table over these local objects only.  StateSet indices stay
    global (StateSets without local drawables get empty regions); matrices are the global scene's; drawable index d of
    the slice is position first + d of the whole list.  Needs host matrices (the big device-synthesised shapes use
    config3_shard)."""
    b = scene.gen["build"]
    sl = slice(first, first + count)
    dg, dm = b["drawable_geom"][sl], b["drawable_ml"][sl]
    geoms, g_inv = np.unique(dg, return_inverse=True)
    lists, l_inv = np.unique(dm, return_inverse=True)
    if isinstance(b["geometries"], dict):
        geo = dict(b["geometries"]); geo["count"] = len(geoms)
    else:
        geo = [b["geometries"][int(g)] for g in geoms]
    if scene.matrices is None:
        raise ValueError("slice_scene needs host matrices")
    starts = np.concatenate([[0], np.cumsum(b["ml_count"].astype(np.int64))])
    rows = np.concatenate([np.arange(starts[k], starts[k + 1]) for k in lists]) if len(lists) else np.zeros(0, np.int64)
    dd = b["drawable_data"]
    out = build_scene(f"{scene.name}[{first}:{first + count}]", geometries=geo, ml_count=b["ml_count"][lists],
                      drawable_geom=g_inv.reshape(-1), drawable_ml=l_inv.reshape(-1), drawable_ps_offset=b["drawable_ps_offset"][sl],
                      state_set=b["state_set"][sl], sphere=b["sphere"][sl], lod_count=b["lod_count"][sl],
                      lod_ps_offset=b["lod_ps_offset"][sl], lod_threshold=b["lod_threshold"][sl],
                      matrices=scene.matrices[rows.astype(np.int64)], drawable_data=None if dd is None else np.asarray(dd)[sl],
                      first_handle=b["first_handle"], force_level=b["force_level"], seed=scene.seed,
                      num_state_sets=b["num_state_sets"])
    out.gen["first"] = int(first)
    return out


# ---------------------------------------------------------------------------------------------------
# matrix synthesis (host).  The torch versions in cadr_b200/frame.py use the same counter-based PRNG.
# ---------------------------------------------------------------------------------------------------
def trs_matrices(pos: np.ndarray, quat: np.ndarray | None, scale: np.ndarray) -> np.ndarray:
    """Column-major mat4 = T * R * S, float32 [m,16]."""
    m = pos.shape[0]
    out = np.zeros((m, 16), dtype=np.float32)
    if quat is None:
        out[:, 0] = scale; out[:, 5] = scale; out[:, 10] = scale
    else:
        x, y, z, w = (quat[:, i].astype(np.float32) for i in range(4))
        s = scale.astype(np.float32)
        out[:, 0] = (1 - 2 * (y * y + z * z)) * s; out[:, 1] = (2 * (x * y + z * w)) * s; out[:, 2] = (2 * (x * z - y * w)) * s
        out[:, 4] = (2 * (x * y - z * w)) * s; out[:, 5] = (1 - 2 * (x * x + z * z)) * s; out[:, 6] = (2 * (y * z + x * w)) * s
        out[:, 8] = (2 * (x * z + y * w)) * s; out[:, 9] = (2 * (y * z - x * w)) * s; out[:, 10] = (1 - 2 * (x * x + y * y)) * s
    out[:, 12:15] = pos.astype(np.float32)
    out[:, 15] = 1
    return out


def uniform_quat(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    u1, u2, u3 = u01(seed, stream, idx), u01(seed, stream + 1, idx), u01(seed, stream + 2, idx)
    a, b = np.sqrt(1 - u1), np.sqrt(u1)
    t2, t3 = np.float32(2 * math.pi) * u2, np.float32(2 * math.pi) * u3
    return np.stack([a * np.sin(t2), a * np.cos(t2), b * np.sin(t3), b * np.cos(t3)], axis=1).astype(np.float32)


def gauss(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    u1 = np.maximum(u01(seed, stream, idx), np.float32(1e-7))
    u2 = u01(seed, stream + 1, idx)
    return (np.sqrt(-2 * np.log(u1)) * np.cos(np.float32(2 * math.pi) * u2)).astype(np.float32)


BOX_SPHERE_RADIUS = math.sqrt(3.0) * 2.43   # box of side 4.86 (Tests.cpp:495-503 with boxSize 4.86)
LOD_PRIMITIVE_SETS = np.array([[36, 0], [24, 36], [12, 60]], dtype=np.uint32)  # Appendix D cfg 3


_BOX_TRIANGLES = np.array([0, 2, 1, 1, 2, 3, 0, 1, 4, 4, 1, 5, 0, 4, 2, 2, 4, 6, 4, 5, 6, 6, 5, 7, 2, 6, 3, 3, 6, 7, 1, 3, 5, 5, 3, 7], np.uint32)


def _box_geometry(lods: bool) -> dict:
    """A real box of side 4.86 (Tests.cpp:495-503): 8 corners x 12 bytes, 36 indices; with LODs 24 and 12 more (the
    first faces again), matching LOD_PRIMITIVE_SETS.  Real so that the consumer-side walk can dereference it."""
    ps = LOD_PRIMITIVE_SETS if lods else LOD_PRIMITIVE_SETS[:1]
    corners = np.array([[(2.43 if v & 1 else -2.43), (2.43 if v & 2 else -2.43), (2.43 if v & 4 else -2.43)] for v in range(8)], np.float32)
    idx = np.concatenate([_BOX_TRIANGLES, _BOX_TRIANGLES[:24], _BOX_TRIANGLES[:12]]) if lods else _BOX_TRIANGLES
    return dict(vertex_bytes=8 * 12, index_bytes=72 * 4 if lods else 36 * 4, primitive_sets=ps, vertices=corners, indices=idx)


def config2(num_drawables: int = 10_000_000, seed: int = 0xC0FFEE02, host_matrices: bool = True,
            cube: float = 2000.0) -> Scene:
    """cfg 2: `num_drawables` drawables x 1 matrix, ONE shared geometry, single StateSet (Appendix D)."""
    n = num_drawables
    mats = None
    if host_matrices:
        i = np.arange(n, dtype=np.uint64)
        pos = np.stack([(u01(seed, s, i) - np.float32(0.5)) * np.float32(cube) for s in (0, 1, 2)], axis=1)
        scale = np.float32(0.5) + np.float32(1.5) * u01(seed, 3, i)
        mats = trs_matrices(pos, None, scale)
    geo = dict(count=1, **_box_geometry(False))
    sphere = np.tile(np.array([0, 0, 0, BOX_SPHERE_RADIUS], dtype=np.float32), (n, 1))
    return build_scene(f"C2:{n}x1", geometries=geo, ml_count=np.ones(n, np.uint32),
                       drawable_geom=np.zeros(n, np.int64), drawable_ml=np.arange(n), drawable_ps_offset=np.zeros(n, np.uint64),
                       state_set=np.zeros(n, np.uint32), sphere=sphere, lod_count=np.ones(n, np.uint32),
                       lod_ps_offset=np.zeros((n, 3), np.uint32), lod_threshold=np.zeros((n, 2), np.float32),
                       matrices=mats, seed=seed, gen=dict(kind="c2", cube=cube))


def config3(num_drawables: int = 100_000, instances: int = 1000, state_sets: int = 64, seed: int = 0xC0FFEE03,
            host_matrices: bool = True, cube: float = 4000.0, sigma: float = 20.0) -> Scene:
    """cfg 3: `num_drawables` geometries x `instances`-matrix lists, 64 StateSets, 3 LODs (Appendix D).
    Drawable d belongs to StateSet d mod 64; the flattened list groups StateSets contiguously."""
    n = num_drawables
    orig, ss = config3_flatten(n, state_sets)          # flatten order -> original id, StateSet
    mats = None
    if host_matrices:
        mats = config3_matrices(seed, np.arange(n, dtype=np.uint64), instances, cube, sigma)
    geo = dict(count=n, **_box_geometry(True))
    sphere = np.tile(np.array([0, 0, 0, BOX_SPHERE_RADIUS], dtype=np.float32), (n, 1))
    return build_scene(f"C3:{n}x{instances}", geometries=geo, ml_count=np.full(n, instances, np.uint32),
                       drawable_geom=orig, drawable_ml=orig, drawable_ps_offset=np.zeros(n, np.uint64),
                       state_set=ss, sphere=sphere, lod_count=np.full(n, 3, np.uint32),
                       lod_ps_offset=np.tile(np.array([0, 8, 16], np.uint32), (n, 1)),
                       lod_threshold=np.tile(np.array([300, 900], np.float32), (n, 1)),
                       matrices=mats, seed=seed, gen=dict(kind="c3", cube=cube, sigma=sigma, instances=instances))


def config3_flatten(num_drawables: int, state_sets: int) -> tuple[np.ndarray, np.ndarray]:
    """Flatten order of cfg 3: position f of the flattened list holds the drawable with original id orig[f] (= its
    geometry = its matrix list) and belongs to StateSet ss[f]; StateSets are contiguous (StateSet.cpp:233-264)."""
    n = num_drawables
    orig = np.concatenate([np.arange(s, n, state_sets) for s in range(state_sets)])
    ss = np.concatenate([np.full(len(range(s, n, state_sets)), s, np.uint32) for s in range(state_sets)])
    return orig, ss


def config3_shard(num_drawables: int, first: int, count: int, instances: int = 1000, state_sets: int = 64,
                  seed: int = 0xC0FFEE03, host_matrices: bool = True, cube: float = 4000.0, sigma: float = 20.0) -> Scene:
    """One rank's part of ONE cfg 3 scene of `num_drawables` drawables: positions [first, first + count) of
    the flattened list, their geometries and matrix lists, and a handle table over these local objects only.  The
    matrices are those of the GLOBAL scene (the PRNG is keyed by the global list id), StateSet indices are global
    (regions of StateSets without local drawables have capacity 0), drawable index d of the shard is position
    first + d of the whole list: the union of the shards' results is the result of config3(num_drawables, ...)."""
    orig, ss_all = config3_flatten(num_drawables, state_sets)
    ids = orig[first:first + count].astype(np.uint64)          # global list / geometry id of each local drawable
    ss = ss_all[first:first + count]
    n = len(ids)
    mats = config3_matrices(seed, ids, instances, cube, sigma) if host_matrices else None
    geo = dict(count=n, **_box_geometry(True))
    sphere = np.tile(np.array([0, 0, 0, BOX_SPHERE_RADIUS], dtype=np.float32), (n, 1))
    sc = build_scene(f"C3:{num_drawables}x{instances}[{first}:{first + n}]", geometries=geo, ml_count=np.full(n, instances, np.uint32),
                     drawable_geom=np.arange(n), drawable_ml=np.arange(n), drawable_ps_offset=np.zeros(n, np.uint64),
                     state_set=ss, sphere=sphere, lod_count=np.full(n, 3, np.uint32),
                     lod_ps_offset=np.tile(np.array([0, 8, 16], np.uint32), (n, 1)),
                     lod_threshold=np.tile(np.array([300, 900], np.float32), (n, 1)),
                     matrices=mats, seed=seed,
                     gen=dict(kind="c3", cube=cube, sigma=sigma, instances=instances, list_ids=ids.astype(np.int64),
                              first=int(first), global_drawables=int(num_drawables)),
                     num_state_sets=state_sets)
    return sc


def config3_matrices(seed: int, lists: np.ndarray, instances: int, cube: float, sigma: float) -> np.ndarray:
    """Matrices of the given matrix lists (cfg 3 recipe): cluster centre uniform in the cube, instances
    Gaussian around it, uniform random rotation, scale U[0.5, 2]."""
    k = np.repeat(lists.astype(np.uint64), instances)
    j = np.tile(np.arange(instances, dtype=np.uint64), len(lists))
    g = k * np.uint64(instances) + j
    centre = np.stack([(u01(seed, s, k) - np.float32(0.5)) * np.float32(cube) for s in (0, 1, 2)], axis=1)
    pos = centre + np.float32(sigma) * np.stack([gauss(seed, 10 + 2 * s, g) for s in (0, 1, 2)], axis=1)
    q = uniform_quat(seed, 20, g)
    scale = np.float32(0.5) + np.float32(1.5) * u01(seed, 30, g)
    return trs_matrices(pos, q, scale)


def config1(boxes_per_side: int = 100, seed: int = 1) -> Scene:
    """cfg 1: RenderingPerformance IndependentBoxesScene — one Geometry and one 1-matrix list per box,
    translate-only matrices on a regular grid (Tests.cpp:456-618).  Tier R only in the reference."""
    m = boxes_per_side
    n = m ** 3
    dist = np.float32(9.72)
    origin = -dist * np.float32(m - 1) / 2
    idx = np.arange(n)
    i, j, k = idx % m, (idx // m) % m, idx // (m * m)
    pos = np.stack([origin + i * dist, origin + j * dist, origin + k * dist], axis=1).astype(np.float32)
    mats = trs_matrices(pos, None, np.ones(n, np.float32))
    geo = dict(count=n, **_box_geometry(False))
    sphere = np.tile(np.array([0, 0, 0, BOX_SPHERE_RADIUS], dtype=np.float32), (n, 1))
    return build_scene(f"C1:{m}^3", geometries=geo, ml_count=np.ones(n, np.uint32), drawable_geom=idx, drawable_ml=idx,
                       drawable_ps_offset=np.zeros(n, np.uint64), state_set=np.zeros(n, np.uint32), sphere=sphere,
                       lod_count=np.ones(n, np.uint32), lod_ps_offset=np.zeros((n, 3), np.uint32),
                       lod_threshold=np.zeros((n, 2), np.float32), matrices=mats, seed=seed, gen=dict(kind="c1"))


def config1_instanced(boxes_per_side: int = 100, seed: int = 1) -> Scene:
    """RenderingPerformance InstancedBoxesScene (Tests.cpp:863-880, generateBoxesScene with instanced = singleGeometry =
    true): ONE drawable, one geometry, one MatrixList holding every box transform (100^3 = 1 M matrices) — the
    instancing extreme next to config1.  The matrices are the same grid as config1's."""
    m = boxes_per_side
    n = m ** 3
    dist = np.float32(9.72)
    origin = -dist * np.float32(m - 1) / 2
    idx = np.arange(n)
    i, j, k = idx % m, (idx // m) % m, idx // (m * m)
    pos = np.stack([origin + i * dist, origin + j * dist, origin + k * dist], axis=1).astype(np.float32)
    mats = trs_matrices(pos, None, np.ones(n, np.float32))
    geo = dict(count=1, **_box_geometry(False))
    sphere = np.array([[0, 0, 0, BOX_SPHERE_RADIUS]], dtype=np.float32)
    return build_scene(f"C1-instanced:{m}^3", geometries=geo, ml_count=np.array([n], np.uint32), drawable_geom=np.zeros(1, np.int64),
                       drawable_ml=np.zeros(1, np.int64), drawable_ps_offset=np.zeros(1, np.uint64), state_set=np.zeros(1, np.uint32),
                       sphere=sphere, lod_count=np.ones(1, np.uint32), lod_ps_offset=np.zeros((1, 3), np.uint32),
                       lod_threshold=np.zeros((1, 2), np.float32), matrices=mats, seed=seed, gen=dict(kind="c1i"))


def reference_camera(frame: int):
    """The camera of examples/RenderingPerformance (main.cpp:936-946,1416-1421): orthographic LH, depth 0..1,
    l/r = -/+960, b/t = -/+540, near 0.5, far 100, eye (x, 0, -50) looking at (x, 0, 0), up (0, -1, 0), x alternating
    1, 0 with the frame number.  -> (planes, eye)"""
    x = float(frame & 1)
    eye = np.array([x, 0.0, -50.0])
    view = look_at_lh(eye, (x, 0.0, 0.0), (0.0, -1.0, 0.0))
    proj = ortho_lh_zo(-960.0, 960.0, -540.0, 540.0, 0.5, 100.0)
    return frustum_planes(proj, view), eye.astype(np.float32)


def random_scene(seed: int, n: int = 300, num_geometries: int = 7, num_lists: int = 50, max_count: int = 70,
                 state_sets: int = 5, first_handle: int = 1, force_level: int = 0, big_lists: int = 0,
                 with_drawable_data: bool = True, cube: float = 400.0, valid_geometry: bool = False,
                 list_counts=None) -> Scene:
    """Ragged scene for parity tests: empty lists, shared lists/geometries, 1-3 LODs, empty spheres, several
    primitive sets per geometry, optional per-drawable data, StateSets of uneven size.  `list_counts` fixes the
    length of every MatrixList (boundary cases of the small / medium / long-list kernels); drawable k then uses
    list k mod len(list_counts) so that every list is referenced."""
    rng = np.random.default_rng(seed)
    geos = []
    for g in range(num_geometries):
        P = int(rng.integers(1, 5))
        if valid_geometry:
            # real geometry for the consumer-side check: 12-byte positions, u32 indices < vertex count, primitive
            # sets that stay inside the index array
            V, I = int(rng.integers(3, 40)), int(rng.integers(6, 120))
            first = rng.integers(0, I - 2, P)
            count = np.array([int(rng.integers(1, min(40, I - f) + 1)) for f in first])
            geos.append(dict(vertices=rng.random((V, 3), dtype=np.float32), indices=rng.integers(0, V, I).astype(np.uint32),
                             primitive_sets=np.stack([count, first], axis=1).astype(np.uint32)))
            continue
        geos.append(dict(vertices=rng.integers(0, 256, int(rng.integers(1, 200)), dtype=np.uint8),
                         indices=rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8),
                         primitive_sets=rng.integers(0, 1 << 20, (P, 2), dtype=np.uint32)))
    counts = rng.integers(0, max_count + 1, num_lists).astype(np.uint32)
    counts[rng.integers(0, num_lists, max(1, num_lists // 10))] = 0            # some empty lists
    counts[rng.integers(0, num_lists, max(1, num_lists // 5))] = 1             # many single-matrix lists
    for b in range(big_lists):
        counts[int(rng.integers(0, num_lists))] = int(rng.integers(1025, 3500))  # lists spanning several work items
    if list_counts is not None:
        counts = np.asarray(list_counts, dtype=np.uint32)
        num_lists = len(counts)
    total = int(counts.sum())
    pos = (rng.random((total, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(cube)
    q = rng.normal(size=(total, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    mats = trs_matrices(pos, q, (0.25 + 3 * rng.random(total)).astype(np.float32))
    dg = rng.integers(0, num_geometries, n)
    dm = rng.integers(0, num_lists, n) if list_counts is None else np.arange(n) % num_lists
    nps = np.array([g["primitive_sets"].shape[0] for g in geos])
    pso = (rng.integers(0, 4, n) % nps[dg]) * 8
    lodc = rng.integers(1, 4, n).astype(np.uint32)
    lpo = ((rng.integers(0, 4, (n, 3)) % nps[dg][:, None]) * 8).astype(np.uint32)
    thr = np.sort(rng.random((n, 2), dtype=np.float32) * np.float32(cube), axis=1)
    sph = np.concatenate([(rng.random((n, 3), dtype=np.float32) - np.float32(0.5)) * 4,
                          (rng.random((n, 1), dtype=np.float32) * 30)], axis=1).astype(np.float32)
    sph[rng.integers(0, n, max(1, n // 20)), 3] = -np.inf                        # empty spheres
    sph[rng.integers(0, n, max(1, n // 40)), 3] = -1.0
    # StateSets: contiguous ranges of uneven size
    cuts = np.sort(rng.integers(0, n + 1, state_sets - 1))
    ss = np.searchsorted(cuts, np.arange(n), side="right").astype(np.uint32)
    dd = rng.random(n) < 0.3 if with_drawable_data else None
    return build_scene(f"random:{seed}", geometries=geos, ml_count=counts, drawable_geom=dg, drawable_ml=dm,
                       drawable_ps_offset=pso.astype(np.uint64), state_set=ss, sphere=sph, lod_count=lodc,
                       lod_ps_offset=lpo, lod_threshold=thr, matrices=mats, drawable_data=dd,
                       first_handle=first_handle, force_level=force_level, seed=seed)
