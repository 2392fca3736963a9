#!/usr/bin/env python
"""Generate tests/golden/ from the REFERENCE ITSELF (run in the build container only; needs /root/reference).

  process_drawables_*.npz   inputs (arena image, drawable list) and the outputs obtained by executing the
                            reference's own compute shader: processDrawables.comp compiled unmodified by the
                            reference's vendored glslangValidator (oracle/Makefile) and interpreted by
                            oracle/spirv_run.py, dispatched as Renderer::recordDrawableProcessing does
                            (Renderer.cpp:684-692, incl. the DispatchBase tail above 32768 drawables).
  allocator_kat.json.gz     command streams and the placements chosen by the reference's own
                            CircularAllocationMemory.h (oracle/ref_alloc_probe.cpp): the scenarios of
                            tests/DataAllocationTest.cpp:62-323 plus random alloc/free sequences.
  parent_child_kat.json.gz  command streams over StateSet-like nodes and the child / parent orders the reference's own
                            ParentChildList.h leaves behind (oracle/ref_parent_child_probe.cpp): the scenario of
                            tests/ParentChildTest.cpp:12-38, repeated links between the same two nodes, removal from
                            either side, clears, and random streams.
  bounding_sphere_ref.npz   (matrix, model-space sphere) pairs and the world-space spheres computed by the reference's
                            own `operator*(const glm::mat4&, BoundingSphere)` (BoundingSphere.h:70-87 with the vendored
                            GLM, oracle/ref_sphere_probe.cpp): pins the sphere transform of the culling extension.

The committed vectors pin oracle/cadr_oracle.c (Tier R) and cadr_b200/host's allocator; the tests never need
/root/reference.
"""
from __future__ import annotations

import gzip
import json
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cadr_b200 import synth  # noqa: E402
from oracle import spirv_run as sr  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")
BASE, LIST, IND, PTR = 0x7F1200000000, 0x7F2000000000, 0x7F3000000000, 0x7F4000000000

SCENES = {
    "L1_ragged": dict(seed=101, n=300),
    "L2_transition_2048": dict(seed=102, n=300, first_handle=1990),
    "L3_forced": dict(seed=103, n=200, first_handle=3000, force_level=3),
    "L3_transition_4194304": dict(seed=104, n=300, first_handle=4_194_250, big_lists=2),
    "dispatch_base_tail": dict(seed=105, n=32768 + 77, num_geometries=9, num_lists=64, max_count=5, state_sets=3),
}


def reloc_words(sc: synth.Scene) -> np.ndarray:
    w = [(node_off // 8 + idx.astype(np.int64))[target != 0] for node_off, idx, target in sc.tables]
    return np.unique(np.concatenate(w)).astype(np.uint32)


def golden_process_drawables():
    for name, kw in SCENES.items():
        sc = synth.random_scene(**kw)
        img = sc.image(BASE)
        dl = np.ascontiguousarray(sc.drawables).copy()
        ind = np.zeros((sc.n, 4), np.uint32)
        ptr = np.zeros((sc.n, 4), np.uint64)
        mem = sr.Memory([(BASE, img), (LIST, dl), (IND, ind), (PTR, ptr)])
        mod = sr.load_module(os.path.join(REF, f"processDrawables_L{sc.handle_level}.spv"))
        push = struct.pack("<4Q", BASE + sc.root_off, LIST, IND, PTR)   # Renderer.cpp:677-682
        sr.dispatch(mod, mem, push, sc.n)
        out = os.path.join(GOLD, f"process_drawables_{name}.npz")
        np.savez_compressed(out, level=sc.handle_level, root_off=sc.root_off, base=BASE, image=img,
                            reloc=reloc_words(sc), drawables=dl, indirect=ind, pointers=ptr,
                            generator=json.dumps(dict(scene="cadr_b200.synth.random_scene", kwargs=kw)))
        print(f"{out}: {sc.n} drawables, level {sc.handle_level}, {os.path.getsize(out) / 1024:.0f} KiB")


def probe(buffer_bytes: int, cmds: list[tuple]) -> list[list[int]]:
    text = f"{buffer_bytes}\n" + "\n".join(" ".join(str(x) for x in c) for c in cmds) + "\n"
    r = subprocess.run([os.path.join(REF, "alloc_probe")], input=text, capture_output=True, text=True, check=True)
    out = []
    for line in r.stdout.strip().splitlines():
        f = line.split()
        out.append([int(x) for x in f[1:]])
    assert len(out) == len(cmds)
    return out


def golden_allocator():
    cases = []

    def add(name, buffer_bytes, cmds):
        cases.append(dict(name=name, buffer=buffer_bytes, cmds=[list(c) for c in cmds], expect=probe(buffer_bytes, cmds)))

    # DataAllocationTest.cpp:87-128: two allocations of size s, both release orders
    for s in list(range(1, 260)):
        add(f"two_of_{s}_fifo", 65536, [("a", 1, s), ("a", 2, s), ("f", 1), ("f", 2)])
        add(f"two_of_{s}_lifo", 65536, [("a", 1, s), ("a", 2, s), ("f", 2), ("f", 1)])
    # DataAllocationTest.cpp:196-323: n allocations of size s within 64 KiB, six release orders + block-2 wrap
    rng = np.random.default_rng(7)
    for s in (1, 16, 17, 48, 64, 65, 100, 128, 129, 259):
        off = 16 if s <= 16 else 32 if s <= 32 else 48 if s <= 48 else 64 if s <= 64 else 128 if s <= 128 else (s + 63) & ~63
        for n in (1, 2, 3, 199, 200, 201, 202, 401, 1029):
            if off * n >= 65536:
                continue
            alloc = [("a", i, s) for i in range(n)]
            orders = {
                "fifo": list(range(n)), "lifo": list(range(n))[::-1],
                "even_odd": list(range(0, n, 2)) + list(range(1, n, 2)),
                "odd_even": list(range(1, n, 2)) + list(range(0, n, 2)),
                "rev_a": list(range(n - 2, -1, -2)) + list(range(n - 1, -1, -2)),
                "rev_b": list(range(n - 1, -1, -2)) + list(range(n - 2, -1, -2)),
            }
            for oname, order in orders.items():
                add(f"grid_{s}x{n}_{oname}", 65536, alloc + [("f", i) for i in order])
            # block-2 scenario (:283-321): a big block pushes the first small one to the buffer end, then wraps
            wrap = [("a", 10_000, 65536 - off), ("a", 0, s), ("f", 10_000)] + [("a", i, s) for i in range(1, n)]
            add(f"wrap_{s}x{n}_fifo", 65536, wrap + [("f", i) for i in range(n)])
            add(f"wrap_{s}x{n}_lifo", 65536, wrap + [("f", i) for i in range(n - 1, -1, -1)])
    # random churn: mixed sizes, interleaved frees, buffers that fill up and wrap
    for k in range(40):
        buf = int(rng.choice([4096, 65536, 1 << 20]))
        live, cmds, nid = [], [], 0
        for _ in range(int(rng.integers(50, 600))):
            if live and rng.random() < 0.45:
                i = live.pop(int(rng.integers(0, len(live))) if rng.random() < 0.5 else 0)
                cmds.append(("f", i))
            else:
                sz = int(rng.choice([1, 8, 16, 40, 64, 100, 128, 1000, 4000, 16384, 60000])) if rng.random() < 0.7 else int(rng.integers(1, buf // 8))
                cmds.append(("a", nid, sz))
                live.append(nid)
                nid += 1
        # frees of allocations that failed (-1) are not issued: filter with a dry run
        res = probe(buf, [c for c in cmds if c[0] == "a"])
        failed = {c[1] for c, r in zip([c for c in cmds if c[0] == "a"], res) if r[1] == -1}
        # a failed alloc in the dry run may succeed in the real interleaving; resolve by simulation
        cmds2, dead = [], set()
        for c in cmds:
            if c[0] == "a":
                cmds2.append(c)
                r = probe(buf, cmds2)[-1]
                if r[1] == -1:
                    dead.add(c[1])
            elif c[1] not in dead:
                cmds2.append(c)
        add(f"random_{k}", buf, cmds2)
    out = os.path.join(GOLD, "allocator_kat.json.gz")
    with gzip.open(out, "wt", compresslevel=9) as f:
        json.dump(dict(source="CadR::CircularAllocationMemory via oracle/ref_alloc_probe.cpp", base=0x10000000, cases=cases), f,
                  separators=(",", ":"))
    print(f"{out}: {len(cases)} cases, {os.path.getsize(out) / 1024:.0f} KiB")


def golden_parent_child():
    import random
    cases = []

    def add(name, nodes, cmds):
        text = f"{nodes}\n" + "\n".join(" ".join(str(x) for x in c) + "\ns" for c in cmds) + "\n"
        r = subprocess.run([os.path.join(REF, "parent_child_probe")], input=text, capture_output=True, text=True, check=True)
        states = r.stdout.strip("\n").split("\n")
        assert len(states) == len(cmds)
        cases.append(dict(name=name, nodes=nodes, cmds=[list(c) for c in cmds], states=[s.strip() for s in states]))

    # tests/ParentChildTest.cpp:12-38
    add("reference_test", 3, [("ac", 0, 1), ("rc", 0, 0), ("cc", 0), ("ap", 1, 0), ("rp", 1, 0), ("cp", 1)])
    # the same child twice under one parent, another parent in between; every removal position from either side
    dup = [("ac", 0, 2), ("ac", 1, 2), ("ac", 0, 3), ("ac", 0, 2), ("ap", 2, 1)]
    for k in range(3):
        add(f"duplicate_links_remove_child_{k}", 4, dup + [("rc", 0, k)])
    for k in range(4):
        add(f"duplicate_links_remove_parent_{k}", 4, dup + [("rp", 2, k)])
    add("clear_child_list_with_duplicates", 4, dup + [("cc", 0)])
    add("clear_parent_list_with_duplicates", 4, dup + [("cp", 2)])
    rng = random.Random(0x5C)
    for case in range(120):
        nodes = rng.randint(2, 7)
        cmds = []
        for _ in range(rng.randint(10, 60)):
            op = rng.choice(["ac", "ac", "ap", "ap", "rc", "rp", "cc", "cp"] if case % 3 else ["ac", "ap", "rc", "rp"])
            a = rng.randrange(nodes)
            if op in ("ac", "ap"):
                b = rng.randrange(nodes)
                if a == b:
                    continue                               # a StateSet is never its own child
                cmds.append((op, a, b))
            elif op == "rc":
                cmds.append(("rc?", a))
            elif op == "rp":
                cmds.append(("rp?", a))
            else:
                cmds.append((op, a))
        # resolve the removal positions against the real list lengths by replaying with the probe
        resolved = []
        for c in cmds:
            if c[0] in ("rc?", "rp?"):
                text = f"{nodes}\n" + "\n".join(" ".join(str(x) for x in r) for r in resolved) + "\ns\n"
                out = subprocess.run([os.path.join(REF, "parent_child_probe")], input=text, capture_output=True, text=True, check=True).stdout
                entry = out.split()[c[1]]
                lst = entry.split(":")[1].split(";")[0 if c[0] == "rc?" else 1].split("=")[1]
                n = len([x for x in lst.split(",") if x])
                if n == 0:
                    continue
                resolved.append((c[0][:2], c[1], rng.randrange(n)))
            else:
                resolved.append(c)
        if resolved:
            add(f"random_{case}", nodes, resolved)
    out = os.path.join(GOLD, "parent_child_kat.json.gz")
    with gzip.open(out, "wt") as f:
        json.dump(dict(source="CadR::ChildList / ParentList (src/CadR/ParentChildList.h) via oracle/ref_parent_child_probe.cpp", cases=cases), f,
                  separators=(",", ":"))
    print(f"{out}: {len(cases)} cases, {sum(len(c['cmds']) for c in cases)} commands, {os.path.getsize(out) / 1024:.0f} KiB")


def golden_bounding_spheres():
    """Random affine transforms of the kinds CAD scenes hold (rotation x non-uniform scale + translation, pure
    translations, mirrored and sheared ones, tiny and huge scales) and random spheres incl. empty ones."""
    rng = np.random.default_rng(0xB5)
    n = 4096
    a = rng.normal(size=(n, 3, 3))
    q, _ = np.linalg.qr(a)                                            # random rotations / reflections
    scale = np.exp(rng.uniform(-3, 3, size=(n, 3)))                   # 0.05 .. 20 per axis
    scale[: n // 8] = scale[: n // 8, :1]                             # uniform scales
    lin = q * scale[:, None, :]
    lin[n // 8: n // 4] = np.eye(3)                                   # pure translations (the boxes scenes)
    lin[-n // 8:] += rng.normal(scale=0.3, size=(n // 8, 3, 3))       # shear
    m = np.zeros((n, 4, 4), np.float32)                               # m[i, c, r]: column-major like glm::mat4
    m[:, :3, :3] = np.transpose(lin, (0, 2, 1))
    m[:, 3, :3] = rng.uniform(-2000, 2000, size=(n, 3))
    m[:, 3, 3] = 1
    spheres = np.concatenate([rng.uniform(-50, 50, size=(n, 3)), rng.uniform(0, 30, size=(n, 1))], axis=1).astype(np.float32)
    spheres[::97, 3] = -np.inf                                        # BoundingSphere::empty()
    spheres[5::97, :3] = 0                                            # centred spheres
    rec = np.concatenate([m.reshape(n, 16), spheres], axis=1).astype(np.float32)
    r = subprocess.run([os.path.join(REF, "sphere_probe")], input=rec.tobytes(), capture_output=True, check=True)
    world = np.frombuffer(r.stdout, dtype=np.float32).reshape(n, 4)
    out = os.path.join(GOLD, "bounding_sphere_ref.npz")
    np.savez_compressed(out, matrices=m.reshape(n, 16), spheres=spheres, world=world,
                        source="CadR::operator*(const glm::mat4&, BoundingSphere), src/CadR/BoundingSphere.h:70-87, via oracle/ref_sphere_probe.cpp (-O1 -ffp-contract=off)")
    print(f"{out}: {n} pairs, {os.path.getsize(out) / 1024:.0f} KiB")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if not os.path.isdir(REF):
        raise SystemExit("oracle/_ref missing: run `make -C oracle ref` in the build container first")
    only = sys.argv[1:]
    if not only or "process_drawables" in only:
        golden_process_drawables()
    if not only or "allocator" in only:
        golden_allocator()
    if not only or "parent_child" in only:
        golden_parent_child()
    if not only or "bounding_spheres" in only:
        golden_bounding_spheres()
