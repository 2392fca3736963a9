"""ctypes binding of libcadr_b200.so (include/cadr_b200.h).

This is the binding a Python host uses; the C++ facade (cadr_b200/host) links the same symbols directly.
The library is built in-tree by cadr_b200/build.py (see __graft_entry__.build).  There is no fallback of any
kind: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcadr_b200.so")

OK = 0
E_LOGIC = -1
E_OUT_OF_RESOURCES = -2
E_TIMEOUT = -3
E_CUDA = -4
E_NO_DEVICE = -5
E_OVERFLOW = -6

CULL_STATUS_REGION_OVERFLOW = 1
CULL_STATUS_CHUNK_OVERFLOW = 2
CULL_STATUS_BAD_RANGE_INDEX = 4
CULL_STATUS_EXCHANGE_TIMEOUT = 8
CULL_HEADER_BYTES = 64


class CadrError(RuntimeError):
    """Mirrors CadR::Error (src/CadR/Exceptions.h:13-40)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class LogicError(CadrError):
    pass


class OutOfResources(CadrError):
    pass


class Timeout(CadrError):
    pass


class NoDevice(CadrError):
    pass


_ERR = {E_LOGIC: LogicError, E_OUT_OF_RESOURCES: OutOfResources, E_TIMEOUT: Timeout, E_NO_DEVICE: NoDevice}


class ExchangePull(C.Structure):
    _fields_ = [("world", C.c_uint32), ("rank", C.c_uint32), ("numRanges", C.c_uint32), ("countersBytes", C.c_uint32),
                ("gatheredCounters", C.c_uint64), ("gatheredInst", C.c_uint64), ("instCapacity", C.c_uint64),
                ("includeLocal", C.c_uint32), ("reserved", C.c_uint32),
                ("regions", C.c_uint64 * 8), ("peerInst", C.c_uint64 * 8)]


class CopyRegion(C.Structure):
    _fields_ = [("dstAddr", C.c_uint64), ("srcOffset", C.c_uint64), ("bytes", C.c_uint64)]


class HandlePatch(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("addr", C.c_uint64)]


class CullParams(C.Structure):
    _fields_ = [
        ("handleTableRoot", C.c_uint64),
        ("handleLevel", C.c_uint32),
        ("numDrawables", C.c_uint32),
        ("drawableList", C.c_uint64),
        ("indirectData", C.c_uint64),
        ("drawablePointers", C.c_uint64),
        ("cullData", C.c_uint64),
        ("planes", (C.c_float * 4) * 6),
        ("eye", C.c_float * 4),
        ("numStateSets", C.c_uint32),
        ("reserved0", C.c_uint32),
        ("stateSetRegions", C.c_uint64),
        ("cmdOut", C.c_uint64),
        ("ptrOut", C.c_uint64),
        ("tagOut", C.c_uint64),
        ("instOut", C.c_uint64),
        ("counters", C.c_uint64),
        ("chunkWorkspace", C.c_uint64),
        ("chunkCapacity", C.c_uint32),
        ("reserved1", C.c_uint32),
        ("exchangeWorld", C.c_uint32),
        ("exchangeRank", C.c_uint32),
        ("exchangeCmdCapacity", C.c_uint32),
        ("reserved2", C.c_uint32),
        ("exchangeCmd", C.c_uint64 * 8),
        ("exchangePtr", C.c_uint64 * 8),
        ("exchangeTag", C.c_uint64 * 8),
        ("drawableBounds", C.c_uint64),
        ("addressDelta", C.c_uint64),
    ]


class ExchangeSync(C.Structure):
    _fields_ = [
        ("world", C.c_uint32), ("rank", C.c_uint32),
        ("frameSeq", C.c_uint64),
        ("localCounters", C.c_uint64),
        ("countersBytes", C.c_uint32), ("timeoutMs", C.c_uint32),
        ("peerCounters", C.c_uint64 * 8),
        ("peerFlags", C.c_uint64 * 8),
    ]


# every symbol include/cadr_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "cadr_b200_abi_version": (C.c_int, []),
    "cadr_b200_last_error": (C.c_char_p, []),
    "cadr_b200_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "cadr_b200_create_address_space_only": (C.c_int, [C.POINTER(_P)]),
    "cadr_b200_destroy": (None, [_P]),
    "cadr_b200_device": (C.c_int, [_P]),
    "cadr_b200_sm_count": (C.c_int, [_P]),
    "cadr_b200_stream": (_P, [_P]),
    "cadr_b200_sync": (C.c_int, [_P, _P, C.c_uint64]),
    "cadr_b200_arena_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_uint64)]),
    "cadr_b200_arena_free": (C.c_int, [_P, C.c_uint64]),
    "cadr_b200_host_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "cadr_b200_host_free": (C.c_int, [_P, _P]),
    "cadr_b200_memcpy_h2d": (C.c_int, [_P, C.c_uint64, _P, C.c_size_t, _P]),
    "cadr_b200_memcpy_d2h": (C.c_int, [_P, _P, C.c_uint64, C.c_size_t, _P]),
    "cadr_b200_memset": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_size_t, _P]),
    "cadr_b200_upload": (C.c_int, [_P, C.POINTER(CopyRegion), C.c_uint32, _P, _P]),
    "cadr_b200_scatter_copy": (C.c_int, [_P, C.POINTER(CopyRegion), C.c_uint32, C.c_uint64, _P]),
    "cadr_b200_patch_handles": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.POINTER(HandlePatch), C.c_uint32, _P]),
    "cadr_b200_process_drawables": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "cadr_b200_record_drawable_processing": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "cadr_b200_cull_compact": (C.c_int, [_P, C.POINTER(CullParams), _P]),
    "cadr_b200_process_and_cull": (C.c_int, [_P, C.POINTER(CullParams), _P]),
    "cadr_b200_compute_drawable_bounds": (C.c_int, [_P, C.POINTER(CullParams), C.c_uint64, C.c_uint64, C.c_uint32, _P]),
    "cadr_b200_ipc_export": (C.c_int, [_P, C.c_uint64, C.c_char_p]),
    "cadr_b200_upload_stage": (C.c_int, [_P, C.POINTER(CopyRegion), C.c_uint32, _P, _P, C.POINTER(C.c_uint64)]),
    "cadr_b200_upload_commit": (C.c_int, [_P, C.c_uint64, _P]),
    "cadr_b200_exchange_pull_instances": (C.c_int, [_P, C.POINTER(ExchangePull), _P]),
    "cadr_b200_ipc_export_range": (C.c_int, [_P, C.c_uint64, C.c_char_p, C.POINTER(C.c_uint64)]),
    "cadr_b200_ipc_import": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_uint64)]),
    "cadr_b200_ipc_close": (C.c_int, [_P, C.c_uint64]),
    "cadr_b200_external_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)]),
    "cadr_b200_external_export_fd": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int)]),
    "cadr_b200_external_import_fd": (C.c_int, [_P, C.c_int, C.c_size_t, C.POINTER(C.c_uint64)]),
    "cadr_b200_external_free": (C.c_int, [_P, C.c_uint64]),
    "cadr_b200_exchange_publish": (C.c_int, [_P, C.POINTER(ExchangeSync), _P]),
    "cadr_b200_exchange_wait": (C.c_int, [_P, C.POINTER(ExchangeSync), _P]),
    "cadr_b200_exchange_publish_and_wait": (C.c_int, [_P, C.POINTER(ExchangeSync), _P]),
    "cadr_b200_consume_check": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "cadr_b200_consume_check_culled": (C.c_int, [_P, C.POINTER(CullParams), C.c_uint32, C.c_uint32, C.c_uint64, _P]),
    "cadr_b200_cull_counters_bytes": (C.c_size_t, [C.c_uint32]),
    "cadr_b200_set_profiling": (C.c_int, [_P, C.c_int]),
    "cadr_b200_kernel_times": (C.c_int, [_P, C.POINTER(C.c_float), C.c_uint32]),
    "cadr_b200_launch_count": (C.c_uint64, [_P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libcadr_b200.so (once).  Raises if it was not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(cadr_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        if l.cadr_b200_abi_version() != 6:
            raise ImportError("libcadr_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(code: int) -> None:
    if code != OK:
        msg = lib().cadr_b200_last_error().decode("utf-8", "replace")
        raise _ERR.get(code, CadrError)(code, msg)


class Context:
    """One context per GPU (cadr_b200_create).  `device=None` makes an address-space-only context that can
    hand out addresses for host-logic tests and refuses every compute call."""

    def __init__(self, device: int | None = 0):
        self._l = lib()
        h = _P()
        if device is None:
            check(self._l.cadr_b200_create_address_space_only(C.byref(h)))
        else:
            check(self._l.cadr_b200_create(int(device), C.byref(h)))
        self._h = h

    def close(self) -> None:
        if self._h:
            self._l.cadr_b200_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- context
    @property
    def device(self) -> int:
        return self._l.cadr_b200_device(self._h)

    @property
    def sm_count(self) -> int:
        return self._l.cadr_b200_sm_count(self._h)

    @property
    def stream(self) -> int:
        return self._l.cadr_b200_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return self._l.cadr_b200_launch_count(self._h)

    def sync(self, stream: int = 0, timeout_ns: int = 0) -> None:
        check(self._l.cadr_b200_sync(self._h, _P(stream), timeout_ns))

    # -- memory
    def arena_alloc(self, nbytes: int) -> int:
        a = C.c_uint64()
        check(self._l.cadr_b200_arena_alloc(self._h, nbytes, C.byref(a)))
        return a.value

    def arena_free(self, addr: int) -> None:
        check(self._l.cadr_b200_arena_free(self._h, addr))

    def host_alloc(self, nbytes: int) -> int:
        p = _P()
        check(self._l.cadr_b200_host_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr: int) -> None:
        check(self._l.cadr_b200_host_free(self._h, _P(ptr)))

    def memcpy_h2d(self, dst: int, src, nbytes: int | None = None, stream: int = 0) -> None:
        ptr, n = _buf(src, nbytes)
        check(self._l.cadr_b200_memcpy_h2d(self._h, dst, ptr, n, _P(stream)))

    def memcpy_d2h(self, dst, src: int, nbytes: int | None = None, stream: int = 0) -> None:
        ptr, n = _buf(dst, nbytes)
        check(self._l.cadr_b200_memcpy_d2h(self._h, ptr, src, n, _P(stream)))

    def memset(self, dst: int, value: int, nbytes: int, stream: int = 0) -> None:
        check(self._l.cadr_b200_memset(self._h, dst, value, nbytes, _P(stream)))

    # -- upload path
    def upload_stage(self, regions, staging, copy_stream: int = 0) -> int:
        """Phase 1 of the two-phase upload: bytes cross PCIe into a device-side slot on `copy_stream` -> ticket."""
        arr, n = _regions(regions)
        ptr, _ = _buf(staging, None)
        t = C.c_uint64()
        check(self._l.cadr_b200_upload_stage(self._h, arr, n, ptr, _P(copy_stream), C.byref(t)))
        return t.value

    def upload_commit(self, ticket: int, stream: int = 0) -> None:
        """Phase 2: `stream` waits for the staging and one scatter launch places the bytes."""
        check(self._l.cadr_b200_upload_commit(self._h, ticket, _P(stream)))

    def upload(self, regions, staging, stream: int = 0) -> None:
        arr, n = _regions(regions)
        ptr, _ = _buf(staging, None)
        check(self._l.cadr_b200_upload(self._h, arr, n, ptr, _P(stream)))

    def scatter_copy(self, regions, staging_dev_addr: int, stream: int = 0) -> None:
        arr, n = _regions(regions)
        check(self._l.cadr_b200_scatter_copy(self._h, arr, n, staging_dev_addr, _P(stream)))

    def patch_handles(self, root: int, level: int, patches, stream: int = 0) -> None:
        n = len(patches)
        arr = (HandlePatch * max(n, 1))()
        for i, (h, a) in enumerate(patches):
            arr[i].handle, arr[i].addr = int(h), int(a)
        check(self._l.cadr_b200_patch_handles(self._h, root, level, arr, n, _P(stream)))

    # -- drawable processing
    def process_drawables(self, root: int, level: int, drawable_list: int, indirect_out: int, pointers_out: int,
                          n: int, stream: int = 0) -> None:
        check(self._l.cadr_b200_process_drawables(self._h, root, level, drawable_list, indirect_out, pointers_out,
                                                  n, _P(stream)))

    def record_drawable_processing(self, host_list, root: int, level: int, drawable_list: int, indirect_out: int,
                                   pointers_out: int, n: int, stream: int = 0) -> None:
        ptr, _ = _buf(host_list, None)
        check(self._l.cadr_b200_record_drawable_processing(self._h, ptr, root, level, drawable_list, indirect_out,
                                                           pointers_out, n, _P(stream)))

    def cull_compact(self, params: CullParams, stream: int = 0) -> None:
        check(self._l.cadr_b200_cull_compact(self._h, C.byref(params), _P(stream)))

    def process_and_cull(self, params: CullParams, stream: int = 0) -> None:
        check(self._l.cadr_b200_process_and_cull(self._h, C.byref(params), _P(stream)))

    def compute_drawable_bounds(self, params: CullParams, bounds_out: int, count: int, indices: int = 0, stream: int = 0) -> None:
        check(self._l.cadr_b200_compute_drawable_bounds(self._h, C.byref(params), bounds_out, indices, count, _P(stream)))

    # -- consumer-side contract check
    def consume_check(self, indirect: int, pointers: int, first: int, n: int, digest_out: int, stream: int = 0) -> None:
        check(self._l.cadr_b200_consume_check(self._h, indirect, pointers, first, n, digest_out, _P(stream)))

    def consume_check_culled(self, params: "CullParams", rng: int, max_commands: int, digest_out: int, stream: int = 0) -> None:
        check(self._l.cadr_b200_consume_check_culled(self._h, C.byref(params), rng, max_commands, digest_out, _P(stream)))

    # -- multi-GPU plumbing
    def ipc_export(self, addr: int) -> bytes:
        buf = C.create_string_buffer(64)
        check(self._l.cadr_b200_ipc_export(self._h, addr, buf))
        return buf.raw

    def exchange_pull_instances(self, pull: "ExchangePull", stream: int = 0) -> None:
        check(self._l.cadr_b200_exchange_pull_instances(self._h, C.byref(pull), _P(stream)))

    def ipc_export_range(self, addr: int) -> tuple[bytes, int]:
        """Any device address inside a cudaMalloc allocation -> (handle of that allocation, offset of addr inside it)."""
        buf, off = C.create_string_buffer(64), C.c_uint64()
        check(self._l.cadr_b200_ipc_export_range(self._h, addr, buf, C.byref(off)))
        return buf.raw, off.value

    def ipc_import(self, handle: bytes) -> int:
        a = C.c_uint64()
        check(self._l.cadr_b200_ipc_import(self._h, handle, C.byref(a)))
        return a.value

    def ipc_close(self, addr: int) -> None:
        check(self._l.cadr_b200_ipc_close(self._h, addr))

    # -- export to a Vulkan consumer / another process
    def external_alloc(self, nbytes: int) -> tuple[int, int]:
        """-> (device address, allocated bytes): a buffer that can be exported as a POSIX fd."""
        a, n = C.c_uint64(), C.c_size_t()
        check(self._l.cadr_b200_external_alloc(self._h, nbytes, C.byref(a), C.byref(n)))
        return a.value, n.value

    def external_export_fd(self, addr: int) -> int:
        fd = C.c_int(-1)
        check(self._l.cadr_b200_external_export_fd(self._h, addr, C.byref(fd)))
        return fd.value

    def external_import_fd(self, fd: int, allocated_bytes: int) -> int:
        a = C.c_uint64()
        check(self._l.cadr_b200_external_import_fd(self._h, fd, allocated_bytes, C.byref(a)))
        return a.value

    def external_free(self, addr: int) -> None:
        check(self._l.cadr_b200_external_free(self._h, addr))

    def exchange_publish(self, sync: "ExchangeSync", stream: int = 0) -> None:
        check(self._l.cadr_b200_exchange_publish(self._h, C.byref(sync), _P(stream)))

    def exchange_publish_and_wait(self, sync: "ExchangeSync", stream: int = 0) -> None:
        check(self._l.cadr_b200_exchange_publish_and_wait(self._h, C.byref(sync), _P(stream)))

    def exchange_wait(self, sync: "ExchangeSync", stream: int = 0) -> None:
        check(self._l.cadr_b200_exchange_wait(self._h, C.byref(sync), _P(stream)))

    def cull_counters_bytes(self, num_state_sets: int) -> int:
        return self._l.cadr_b200_cull_counters_bytes(num_state_sets)

    # -- timing
    def set_profiling(self, enabled: bool) -> None:
        check(self._l.cadr_b200_set_profiling(self._h, 1 if enabled else 0))

    def kernel_times(self) -> list[float]:
        ms = (C.c_float * 5)()
        check(self._l.cadr_b200_kernel_times(self._h, ms, 5))
        return list(ms)


def _buf(obj, nbytes):
    """(void*, nbytes) of an int address, a numpy array or anything with the buffer protocol."""
    if isinstance(obj, int):
        return _P(obj), (nbytes or 0)
    try:
        import numpy as np
        if isinstance(obj, np.ndarray):
            if not obj.flags["C_CONTIGUOUS"]:
                raise ValueError("numpy buffer must be C-contiguous")
            return _P(obj.ctypes.data), (obj.nbytes if nbytes is None else nbytes)
    except ImportError:
        pass
    mv = memoryview(obj)
    a = (C.c_char * mv.nbytes).from_buffer(obj)
    return C.cast(a, _P), (mv.nbytes if nbytes is None else nbytes)


def _regions(regions):
    """Accepts a list of (dst, srcOffset, bytes) or an (n,3) uint64 numpy array."""
    try:
        import numpy as np
        if isinstance(regions, np.ndarray):
            r = np.ascontiguousarray(regions, dtype=np.uint64).reshape(-1, 3)
            return r.ctypes.data_as(C.POINTER(CopyRegion)), r.shape[0]  # data_as keeps `r` alive
    except ImportError:
        pass
    n = len(regions)
    arr = (CopyRegion * max(n, 1))()
    for i, (d, s, b) in enumerate(regions):
        arr[i].dstAddr, arr[i].srcOffset, arr[i].bytes = int(d), int(s), int(b)
    return arr, n
