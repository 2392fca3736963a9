"""Minimal SPIR-V interpreter — just enough to EXECUTE the reference's own compute shader on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/cadr_oracle.c).  No Vulkan ICD exists in this environment, so
the reference shader /root/reference/src/CadR/shaders/processDrawables.comp cannot run natively.  What does work
is the reference's vendored compiler: `3rdParty/Vulkan/bin/glslangValidator --target-env vulkan1.2
-DHANDLE_LEVEL_n` turns the unmodified GLSL into SPIR-V 1.5.  This module interprets that binary one workgroup
at a time (local_size is 1x1x1, :14), with PhysicalStorageBuffer64 pointers resolved against a segment model
of device memory, so its outputs ARE "outputs of the reference itself, run here".  They pin the C oracle
(tests/golden/, oracle/make_golden.py).

Supported: the integer / pointer subset that shader uses (OpLoad/OpStore/OpAccessChain on Function,
PushConstant, Input and PhysicalStorageBuffer storage, OpIAdd/OpIMul/OpUConvert/OpShiftRightLogical/
OpBitwiseAnd/OpConvertUToPtr/OpConvertPtrToU/OpFunctionCall, ...).  Anything else raises NotImplementedError
rather than guessing.
"""
from __future__ import annotations

import struct

import numpy as np

OP = {
    17: "Capability", 11: "ExtInstImport", 14: "MemoryModel", 15: "EntryPoint", 16: "ExecutionMode", 3: "Source",
    4: "SourceExtension", 5: "Name", 6: "MemberName", 71: "Decorate", 72: "MemberDecorate",
    19: "TypeVoid", 21: "TypeInt", 22: "TypeFloat", 23: "TypeVector", 24: "TypeMatrix", 28: "TypeArray",
    29: "TypeRuntimeArray", 30: "TypeStruct", 32: "TypePointer", 33: "TypeFunction", 39: "TypeForwardPointer",
    43: "Constant", 44: "ConstantComposite", 54: "Function", 55: "FunctionParameter", 56: "FunctionEnd",
    57: "FunctionCall", 59: "Variable", 61: "Load", 62: "Store", 65: "AccessChain", 113: "UConvert",
    114: "SConvert", 120: "ConvertUToPtr", 117: "ConvertPtrToU", 124: "Bitcast", 128: "IAdd", 130: "ISub",
    132: "IMul", 194: "ShiftRightLogical", 196: "ShiftLeftLogical", 197: "BitwiseOr", 199: "BitwiseAnd",
    248: "Label", 253: "Return", 254: "ReturnValue", 8: "Line", 7: "String", 330: "ModuleProcessed",
    4417: "ExtensionKHR", 10: "Extension",
}
SC_INPUT, SC_PUSH, SC_FUNCTION, SC_PSB = 1, 9, 7, 5349
DEC_ARRAY_STRIDE, DEC_BUILTIN, DEC_OFFSET = 6, 11, 35
BUILTIN_WORKGROUP_ID = 26


class Memory:
    """Device memory as segments [(base address, uint8 ndarray)] — the same model the C oracle uses."""

    def __init__(self, segments):
        self.segs = [(int(b), np.ascontiguousarray(a).view(np.uint8).reshape(-1)) for b, a in segments]

    def _find(self, addr, n):
        for base, arr in self.segs:
            if base <= addr and addr + n <= base + arr.nbytes:
                return arr, addr - base
        raise MemoryError(f"access of {n} bytes at 0x{addr:x} is outside device memory (UB in the reference)")

    def load(self, addr, n):
        arr, off = self._find(addr, n)
        return int.from_bytes(arr[off:off + n].tobytes(), "little")

    def store(self, addr, n, value):
        arr, off = self._find(addr, n)
        arr[off:off + n] = np.frombuffer(int(value).to_bytes(n, "little"), dtype=np.uint8)


class Ptr:
    """A logical pointer: Function/PushConstant/Input variables hold Python values, PSB pointers are addresses."""
    __slots__ = ("sc", "var", "path", "addr", "type")

    def __init__(self, sc, type_, var=None, path=(), addr=0):
        self.sc, self.type, self.var, self.path, self.addr = sc, type_, var, path, addr


class Module:
    def __init__(self, words):
        if words[0] != 0x07230203:
            raise ValueError("not a SPIR-V module")
        self.version = words[1]
        self.types, self.consts, self.decor, self.mdecor = {}, {}, {}, {}
        self.globals, self.functions, self.entry = {}, {}, None
        i, cur = 5, None
        while i < len(words):
            wc, op = words[i] >> 16, words[i] & 0xFFFF
            ops = words[i + 1:i + wc]
            name = OP.get(op)
            if name is None:
                raise NotImplementedError(f"SPIR-V opcode {op}")
            if name == "EntryPoint":
                self.entry = ops[1]
            elif name == "Decorate":
                self.decor.setdefault(ops[0], {})[ops[1]] = ops[2:] and ops[2]
            elif name == "MemberDecorate":
                self.mdecor.setdefault(ops[0], {}).setdefault(ops[1], {})[ops[2]] = ops[3:] and ops[3]
            elif name.startswith("Type"):
                self._type(name, ops)
            elif name == "Constant":
                t = self.types[ops[0]]
                v = ops[2] | (ops[3] << 32 if len(ops) > 3 else 0)
                self.consts[ops[1]] = v & ((1 << t["width"]) - 1)
            elif name == "ConstantComposite":
                self.consts[ops[1]] = [self.consts[c] for c in ops[2:]]
            elif name == "Variable" and cur is None:
                self.globals[ops[1]] = (ops[0], ops[2])
            elif name == "Function":
                cur = {"id": ops[1], "params": [], "body": [], "rtype": ops[0]}
                self.functions[ops[1]] = cur
            elif name == "FunctionParameter":
                cur["params"].append(ops[1])
            elif name == "FunctionEnd":
                cur = None
            elif cur is not None:
                cur["body"].append((name, ops))
            i += wc

    def _type(self, name, ops):
        k = name[4:]
        t = {"kind": k}
        if k == "Int":
            t.update(width=ops[1], size=ops[1] // 8)
        elif k == "Float":
            t.update(width=ops[1], size=ops[1] // 8)
        elif k == "Vector":
            t.update(elem=ops[1], count=ops[2], size=self.types[ops[1]]["size"] * ops[2])
        elif k == "Matrix":
            t.update(elem=ops[1], count=ops[2], size=self.types[ops[1]]["size"] * ops[2])
        elif k == "Array":
            t.update(elem=ops[1], length=ops[2])
        elif k == "RuntimeArray":
            t.update(elem=ops[1])
        elif k == "Struct":
            t.update(members=list(ops[1:]))
        elif k == "Pointer":
            t.update(sc=ops[1], pointee=ops[2], size=8)
        elif k == "ForwardPointer":
            return
        self.types[ops[0]] = t


class Invocation:
    def __init__(self, mod: Module, mem: Memory, push: bytes, workgroup_id):
        self.m, self.mem, self.push, self.wg = mod, mem, push, workgroup_id

    # -- typed access to PushConstant / Input / Function storage ---------------------------------------
    def _offset(self, struct_id, member):
        return self.m.mdecor[struct_id][member][DEC_OFFSET]

    def _load_push(self, type_id, off):
        t = self.m.types[type_id]
        if t["kind"] in ("Int", "Pointer"):
            return int.from_bytes(self.push[off:off + t["size"]], "little")
        raise NotImplementedError(f"push-constant load of {t['kind']}")

    def access_chain(self, res_type, base, idx):
        ptype = self.m.types[res_type]
        if isinstance(base, Ptr) and base.sc == SC_PSB or not isinstance(base, Ptr):
            addr = base.addr if isinstance(base, Ptr) else int(base[1])
            tid = base.type if isinstance(base, Ptr) else base[0]
            for ix in idx:
                t = self.m.types[tid]
                if t["kind"] == "Struct":
                    addr += self._offset(tid, ix)
                    tid = t["members"][ix]
                elif t["kind"] in ("Array", "RuntimeArray"):
                    addr += self.m.decor[tid][DEC_ARRAY_STRIDE] * ix
                    tid = t["elem"]
                else:
                    raise NotImplementedError(f"access chain into {t['kind']}")
            return Ptr(SC_PSB, tid, addr=addr & ((1 << 64) - 1))
        return Ptr(base.sc, ptype["pointee"], var=base.var, path=base.path + tuple(idx))

    def load(self, ptr: Ptr):
        t = self.m.types[ptr.type]
        if ptr.sc == SC_PSB:
            if t["kind"] in ("Int", "Pointer"):
                v = self.mem.load(ptr.addr, t["size"])
                return (t["pointee"], v) if t["kind"] == "Pointer" else v
            raise NotImplementedError(f"PSB load of {t['kind']}")
        if ptr.sc == SC_PUSH:
            tid, off = self.m.types[self.m.globals[ptr.var][0]]["pointee"], 0
            for ix in ptr.path:
                off += self._offset(tid, ix)
                tid = self.m.types[tid]["members"][ix]
            v = self._load_push(tid, off)
            tt = self.m.types[tid]
            return (tt["pointee"], v) if tt["kind"] == "Pointer" else v
        if ptr.sc == SC_INPUT:
            if self.m.decor[ptr.var].get(DEC_BUILTIN) != BUILTIN_WORKGROUP_ID:
                raise NotImplementedError("input builtin")
            return self.wg[ptr.path[0]] if ptr.path else list(self.wg)
        if ptr.sc == SC_FUNCTION:
            return self.locals[ptr.var]
        raise NotImplementedError(f"load from storage class {ptr.sc}")

    def store(self, ptr: Ptr, value):
        if ptr.sc == SC_FUNCTION:
            self.locals[ptr.var] = value
            return
        if ptr.sc == SC_PSB:
            t = self.m.types[ptr.type]
            if t["kind"] == "Int":
                self.mem.store(ptr.addr, t["size"], value)
                return
        raise NotImplementedError(f"store to storage class {ptr.sc}")

    # -- execution -------------------------------------------------------------------------------------
    def run(self):
        self.locals = {}
        return self.call(self.m.entry, [])

    def call(self, fid, args):
        f = self.m.functions[fid]
        v = dict(zip(f["params"], args))
        m = self.m

        def val(i):
            if i in v:
                return v[i]
            if i in m.consts:
                return m.consts[i]
            if i in m.globals:
                tid, sc = m.globals[i]
                return Ptr(sc, m.types[tid]["pointee"], var=i)
            raise KeyError(i)

        def width(tid):
            return m.types[tid]["width"]

        for name, o in f["body"]:
            if name in ("Label", "Line"):
                continue
            if name == "Variable":
                v[o[1]] = Ptr(SC_FUNCTION, m.types[o[0]]["pointee"], var=(fid, o[1]))
            elif name == "AccessChain":
                v[o[1]] = self.access_chain(o[0], val(o[2]), [val(i) for i in o[3:]])
            elif name == "Load":
                v[o[1]] = self.load(val(o[2]))
            elif name == "Store":
                self.store(val(o[0]), val(o[1]))
            elif name == "IAdd":
                v[o[1]] = (val(o[2]) + val(o[3])) & ((1 << width(o[0])) - 1)
            elif name == "ISub":
                v[o[1]] = (val(o[2]) - val(o[3])) & ((1 << width(o[0])) - 1)
            elif name == "IMul":
                v[o[1]] = (val(o[2]) * val(o[3])) & ((1 << width(o[0])) - 1)
            elif name == "UConvert":
                v[o[1]] = val(o[2]) & ((1 << width(o[0])) - 1)   # zero-extend or truncate
            elif name == "ShiftRightLogical":
                v[o[1]] = (val(o[2]) >> val(o[3])) & ((1 << width(o[0])) - 1)
            elif name == "ShiftLeftLogical":
                v[o[1]] = (val(o[2]) << val(o[3])) & ((1 << width(o[0])) - 1)
            elif name == "BitwiseAnd":
                v[o[1]] = val(o[2]) & val(o[3])
            elif name == "BitwiseOr":
                v[o[1]] = val(o[2]) | val(o[3])
            elif name == "ConvertUToPtr":
                v[o[1]] = (m.types[o[0]]["pointee"], val(o[2]))
            elif name == "ConvertPtrToU":
                p = val(o[2])
                v[o[1]] = (p.addr if isinstance(p, Ptr) else p[1]) & ((1 << width(o[0])) - 1)
            elif name == "FunctionCall":
                v[o[1]] = self.call(o[2], [val(i) for i in o[3:]])
            elif name == "Return":
                return None
            elif name == "ReturnValue":
                return val(o[0])
            else:
                raise NotImplementedError(f"SPIR-V instruction Op{name}")
        return None


def load_module(path: str) -> Module:
    data = open(path, "rb").read()
    return Module(list(struct.unpack(f"<{len(data) // 4}I", data)))


def dispatch(mod: Module, mem: Memory, push_constants: bytes, num_workgroups: int) -> None:
    """vkCmdDispatch / vkCmdDispatchBase exactly as Renderer::recordDrawableProcessing issues them
    (/root/reference/src/CadR/Renderer.cpp:684-692): grid (<=32768, y) + a DispatchBase tail."""
    if num_workgroups <= 32768:
        groups = [(x, 0, 0) for x in range(num_workgroups)]
    else:
        y = (num_workgroups - 1) // 32768
        x = (num_workgroups - 1) % 32768 + 1
        groups = [(gx, gy, 0) for gy in range(y) for gx in range(32768)] + [(gx, y, 0) for gx in range(x)]
    for g in groups:
        Invocation(mod, mem, push_constants, g).run()
