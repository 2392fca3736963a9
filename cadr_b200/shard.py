"""Multi-GPU sharding of the drawable-processing path: one process per GPU.

  partition()   cut the flattened drawable list into `world` contiguous slices balanced by INSTANCE count, at drawable
                boundaries only; StateSet ranges are contiguous in flatten order (StateSet.cpp:233-264), so
                concatenating per-rank outputs in rank order keeps every StateSet's commands contiguous per rank.
  TierRGather   the fixed-size records of the processing pass (IndirectData 16 B, DrawablePointers 32 B per drawable):
                every rank's slice lands at its global drawable index in one array on every rank, i.e. exactly the
                arrays a single GPU would have written for the whole list (the pointers are addresses of the owning GPU).
  Exchange      the one real exchange step of the culling pass: all-gather of every rank's compacted command list (commands, forwarded
                pointers, tags) and per-range counters into one buffer on every rank, plus a directory
                {rank, range} -> (command offset, count) a renderer walks.  Instance-index lists stay on the owning GPU
                (gathering 4 B per survivor costs more NVLink time than the cull itself at p ~ 0.5).

torch.distributed is plumbing here: NCCL over NVLink on GPUs, gloo on CPU tensors in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition(instance_counts: np.ndarray, world: int) -> list[tuple[int, int]]:
    """-> [(first, count)] per rank; slices are contiguous, cover [0, n), and differ by less than one drawable's
    instances from the ideal share (empty lists weigh one so that drawable-only work is balanced too)."""
    w = np.maximum(np.asarray(instance_counts, dtype=np.int64), 1)
    n = len(w)
    cum = np.concatenate([[0], np.cumsum(w)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        # pick the boundary closer to the target
        if k > cuts[-1] and k <= n and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        cuts.append(max(k, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world)]


class TierRGather:
    """Gathers the per-drawable Tier R records of all ranks into whole-scene arrays ("Tier R (fixed-size
    records)").  Slices are contiguous in flatten order and of different lengths, so rank r's records are broadcast
    straight into rows [first_r, first_r + count_r) of the result: no padding, no re-packing, and the result is
    bit-identical to a single-GPU pass over the whole list.  `slices` = partition(...) of the global list."""

    def __init__(self, slices: list[tuple[int, int]], device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if len(slices) != self.world:
            raise ValueError("TierRGather: one slice per rank expected")
        first = 0
        for f, c in slices:
            if f != first or c < 0:
                raise ValueError("TierRGather: slices must be contiguous and start at 0")
            first += c
        self.slices = [(int(f), int(c)) for f, c in slices]
        self.n = first
        self.indirect = torch.zeros(self.n * 16, dtype=torch.uint8, device=device)     # cadr_indirect_data[n]
        self.pointers = torch.zeros(self.n * 32, dtype=torch.uint8, device=device)     # cadr_drawable_pointers[n]

    def run(self, indirect: torch.Tensor, pointers: torch.Tensor) -> None:
        """`indirect` / `pointers`: this rank's uint8 views of its slice's records (count_r * 16 / * 32 bytes)."""
        f, c = self.slices[self.rank]
        if indirect.numel() < c * 16 or pointers.numel() < c * 32:
            raise ValueError("TierRGather: this rank's record arrays are shorter than its slice")
        self.indirect[f * 16:(f + c) * 16].copy_(indirect[:c * 16])
        self.pointers[f * 32:(f + c) * 32].copy_(pointers[:c * 32])
        for r, (rf, rc) in enumerate(self.slices):
            if rc == 0:
                continue
            src = dist.get_global_rank(self.group, r) if self.group is not None else r
            dist.broadcast(self.indirect[rf * 16:(rf + rc) * 16], src=src, group=self.group)
            dist.broadcast(self.pointers[rf * 32:(rf + rc) * 32], src=src, group=self.group)

    def records(self) -> tuple[np.ndarray, np.ndarray]:
        """-> (indirect [n,4] u32, pointers [n,4] u64) on the host; `owner(i)` tells whose addresses row i holds."""
        return (self.indirect.cpu().numpy().view(np.uint32).reshape(-1, 4), self.pointers.cpu().numpy().view(np.uint64).reshape(-1, 4))

    def owner(self, drawable: int) -> int:
        for r, (f, c) in enumerate(self.slices):
            if f <= drawable < f + c:
                return r
        raise IndexError(drawable)


class Exchange:
    """All-gather of per-rank Tier X outputs.  Buffers are padded to the largest rank's capacity so that one
    all_gather_into_tensor per array suffices (NVSwitch: every peer at full bandwidth; the payload is small)."""

    def __init__(self, cmd_capacity: int, num_ranges: int, device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        caps = torch.tensor([cmd_capacity, num_ranges], dtype=torch.int64, device=device)
        allcaps = torch.empty(self.world * 2, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allcaps, caps, group=group)
        allcaps = allcaps.view(self.world, 2).cpu()
        self.cmd_cap = int(allcaps[:, 0].max())
        self.num_ranges = int(allcaps[:, 1].max())
        self.sizes = dict(cmd=self.cmd_cap * 20, ptr=self.cmd_cap * 32, tag=self.cmd_cap * 8, counters=64 + 8 * self.num_ranges)
        self.gathered = {k: torch.empty(self.world * v, dtype=torch.uint8, device=device) for k, v in self.sizes.items()}
        self.bytes_per_rank = sum(self.sizes.values())

    def run(self, cmd: torch.Tensor, ptr: torch.Tensor, tag: torch.Tensor, counters: torch.Tensor) -> None:
        """Inputs are this rank's uint8 views (at least `sizes[...]` bytes each, or shorter: then they are padded)."""
        for k, t in (("cmd", cmd), ("ptr", ptr), ("tag", tag), ("counters", counters)):
            n = self.sizes[k]
            if t.numel() < n:
                t = torch.cat([t, torch.zeros(n - t.numel(), dtype=torch.uint8, device=t.device)])
            dist.all_gather_into_tensor(self.gathered[k], t[:n].contiguous(), group=self.group)

    def directory(self, regions_per_rank: list[np.ndarray]) -> list[dict]:
        """Host view after a sync: one entry per (rank, range) with commands: where they start in the gathered command
        buffer and how many there are."""
        cnt = self.gathered["counters"].view(self.world, -1)[:, 64:].contiguous().view(torch.int64).cpu().numpy()
        out = []
        for r in range(self.world):
            reg = regions_per_rank[r]
            for s in range(reg.shape[0]):
                c = int(cnt[r, s] & 0xFFFFFFFF)
                if c:
                    out.append(dict(rank=r, range=s, first_command=r * self.cmd_cap + int(reg[s, 0]), count=c,
                                    instances=int(cnt[r, s] >> 32)))
        return out

    def commands(self) -> dict:
        """numpy views of the gathered arrays: cmd [world*cap,5] u32, ptr [.,4] u64, tag [.,2] u32."""
        g = {k: v.cpu().numpy() for k, v in self.gathered.items()}
        return dict(cmd=g["cmd"].view(np.uint32).reshape(-1, 5), ptr=g["ptr"].view(np.uint64).reshape(-1, 4),
                    tag=g["tag"].view(np.uint32).reshape(-1, 2))


class PeerExchange:
    """The fused exchange: no collective call per frame.  Every rank owns gathered arrays (commands, pointers, tags,
    counters) and a flag array, all exported through CUDA IPC and mapped by every peer once at set-up.  During a
    frame the cull kernels store each emitted record directly into ALL ranks' gathered arrays over NVLink
    (cadr_cull_params.exchange*), so the transfer overlaps the cull item by item; afterwards a one-CTA kernel
    publishes the rank's counters and raises a frame flag on every peer, and a second one makes the stream wait for
    all peers' flags.  Gathered arrays are double-buffered by frame parity: a rank can run at most one frame ahead
    of a peer (it cannot pass the wait), so set k&1 is never overwritten while a peer still reads frame k-2... k."""

    def __init__(self, ctx, cmd_capacity: int, num_ranges: int, group=None, *, regions: np.ndarray | None = None,
                 inst_out: int = 0, arena: int = 0, first_drawable: int = 0, sets: int = 2, deferred_wait: bool = False):
        """regions / inst_out / arena (optional, all or none): this rank's region table [S,4] u32, its instance-index
        buffer and the arena that holds its geometry and matrix lists.  They are exported to the peers as well, which
        makes every rank's result CONSUMABLE on any GPU (consume_params): commands and counters are local copies, instance
        indices and matrices are read through the peer mappings on demand (mitigation iii).  inst_out may be a
        pair of buffers: begin_frame then alternates them by frame parity like the gathered arrays, so that a peer may
        still read frame k's runs while this rank already culls frame k + 1 (it cannot get further ahead: it cannot pass
        the wait of frame k + 1 before every peer has published it).  first_drawable = position of this rank's first
        drawable in the whole flattened list (the tag's drawable index is slice-relative).

        deferred_wait (needs sets >= 4): end_frame(k) makes the stream wait for the peers' frame k-1 instead of frame k, and
        does so BEFORE it publishes frame k.  A rank then never idles for the slowest rank of the frame it has just finished
        (the max-over-ranks skew that is the exchange's only real cost): it starts frame k+1 at once and only needs the
        peers to be less than a whole frame behind.  The price is one frame of latency: after end_frame(k) the gathered
        result that is complete on this GPU is frame k-1's (`complete_frame`); finish() waits for the last one.  Why four
        sets: a consumer of frame k-1 is queued behind end_frame(k), i.e. behind publish(k); peer A may write frame k+2 as
        soon as it has seen everybody's publish(k) - while that consumer still runs - and frame k+3 only after everybody's
        publish(k+1), which sits behind the consumer on this stream.  So frames k-1 .. k+2 must not share a set."""
        if deferred_wait and sets < 4:
            raise ValueError("PeerExchange: deferred_wait needs at least four sets of gathered arrays")
        self.nsets, self.deferred = int(sets), bool(deferred_wait)
        self.fused_sync = True             # publish + wait as one kernel (False: two launches, as in round 1)
        self.timeout_ms = 0                # > 0: budget of a stream-side wait for the peers (cadr_exchange_sync.timeoutMs); 0: wait for ever
        import ctypes as C
        from . import _capi
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        caps = [None] * self.world
        dist.all_gather_object(caps, (int(cmd_capacity), int(num_ranges)), group=group)
        self.cmd_cap = max(c[0] for c in caps)
        self.num_ranges = max(c[1] for c in caps)
        self.counters_bytes = 64 + 8 * self.num_ranges
        sizes = dict(cmd=self.world * self.cmd_cap * 20, ptr=self.world * self.cmd_cap * 32, tag=self.world * self.cmd_cap * 8,
                     counters=self.world * self.counters_bytes)
        self.sizes = sizes
        # `sets` sets of gathered arrays (two: frame parity) + one flag array
        self.local = [{k: ctx.arena_alloc(max(v, 256)) for k, v in sizes.items()} for _ in range(self.nsets)]
        self.flags = ctx.arena_alloc(256)
        ctx.memset(self.flags, 0, 256)
        for st in self.local:
            ctx.memset(st["counters"], 0, sizes["counters"])
        ctx.sync()
        mine = dict(sets=[{k: ctx.ipc_export(a) for k, a in st.items()} for st in self.local], flags=ctx.ipc_export(self.flags))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.peer = []   # [rank] -> {"sets": [{cmd,ptr,tag,counters}], "flags": addr}
        self._imported = []
        for r, h in enumerate(everyone):
            if r == self.rank:
                self.peer.append(dict(sets=self.local, flags=self.flags))
                continue
            sets = []
            for st in h["sets"]:
                m = {k: ctx.ipc_import(v) for k, v in st.items()}
                self._imported += list(m.values())
                sets.append(m)
            f = ctx.ipc_import(h["flags"])
            self._imported.append(f)
            self.peer.append(dict(sets=sets, flags=f))
        self.frame = 0
        self._C, self._capi = C, _capi
        # -- what a consumer on another GPU needs besides the gathered commands ------------------------------------
        self.consumable = regions is not None
        if self.consumable:
            if not inst_out or not arena:
                raise ValueError("PeerExchange: regions, inst_out and arena go together")
            reg = np.zeros((self.num_ranges, 4), np.uint32)
            reg[:regions.shape[0]] = regions
            self.inst_sets = [int(a) for a in (inst_out if isinstance(inst_out, (list, tuple)) else [inst_out])]
            mine = dict(regions=reg, first=int(first_drawable), arena=int(arena),
                        inst=[ctx.ipc_export_range(a) for a in self.inst_sets], mem=ctx.ipc_export_range(arena))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            self.first_drawable = [e["first"] for e in everyone]
            self.peer_regions = [e["regions"] for e in everyone]
            self.regions_dev = []                  # [rank] -> device copy of that rank's region table
            self.peer_inst = []                    # [rank] -> its instance-index buffer(s) as mapped here (one, or one per frame parity)
            self.peer_delta = []                   # [rank] -> (its arena as mapped here) - (address on the owner), mod 2^64
            opened = {}                            # one mapping per exported allocation (two buffers may share one)

            def open_(h_off):
                h, off = h_off
                if h not in opened:
                    opened[h] = ctx.ipc_import(h)
                    self._imported.append(opened[h])
                return opened[h] + off

            for r, e in enumerate(everyone):
                a = ctx.arena_alloc(max(reg.nbytes, 256))
                ctx.memcpy_h2d(a, np.ascontiguousarray(e["regions"]))
                self.regions_dev.append(a)
                if r == self.rank:
                    self.peer_inst.append(list(self.inst_sets)); self.peer_delta.append(0)
                else:
                    self.peer_inst.append([open_(x) for x in e["inst"]])
                    self.peer_delta.append((open_(e["mem"]) - e["arena"]) & 0xFFFFFFFFFFFFFFFF)
            self.digests = ctx.arena_alloc(max(16 * self.num_ranges, 256))
            ctx.sync()
        dist.barrier(group=group)

    def begin_frame(self, params) -> None:
        """Point the cull at this frame's gathered arrays (modifies `params` in place)."""
        self.frame += 1
        k = self.frame % self.nsets
        params.exchangeWorld, params.exchangeRank, params.exchangeCmdCapacity = self.world, self.rank, self.cmd_cap
        for r in range(self.world):
            s = self.peer[r]["sets"][k]
            params.exchangeCmd[r], params.exchangePtr[r], params.exchangeTag[r] = s["cmd"], s["ptr"], s["tag"]
        if self.consumable and len(self.inst_sets) > 1:
            params.instOut = self.inst_sets[self.frame % len(self.inst_sets)]

    def inst_of(self, r: int) -> int:
        """Rank r's instance-index buffer of the current frame as mapped on this GPU."""
        bufs = self.peer_inst[r]
        return bufs[self.frame % len(bufs)]

    def _sync(self, local_counters: int, frame: int | None = None):
        frame = self.frame if frame is None else frame
        s = self._capi.ExchangeSync()
        s.world, s.rank, s.frameSeq = self.world, self.rank, frame
        s.localCounters, s.countersBytes, s.timeoutMs = local_counters, self.counters_bytes, int(self.timeout_ms)
        k = frame % self.nsets
        for r in range(self.world):
            s.peerCounters[r] = self.peer[r]["sets"][k]["counters"]
            s.peerFlags[r] = self.peer[r]["flags"]
        return s

    def end_frame(self, local_counters: int, stream: int = 0) -> None:
        """After the cull kernels of the frame: publish counters + flag to every peer and wait for all peers - for this
        frame, or with deferred_wait for the previous one and before the publish (see __init__)."""
        s = self._sync(local_counters)
        if not self.deferred:
            if self.fused_sync:
                self.ctx.exchange_publish_and_wait(s, stream)      # one launch: publish, then spin on the local flags
            else:
                self.ctx.exchange_publish(s, stream)
                self.ctx.exchange_wait(s, stream)
            return
        if self.frame > 1:
            self.ctx.exchange_wait(self._sync(local_counters, self.frame - 1), stream)
        self.ctx.exchange_publish(s, stream)
        self._last_counters = local_counters

    @property
    def complete_frame(self) -> int:
        """The newest frame whose gathered result is complete on this GPU once the stream has passed end_frame()."""
        return self.frame - 1 if self.deferred else self.frame

    def finish(self, stream: int = 0) -> None:
        """deferred_wait: wait for the peers' last frame too (end of a run, or before reading the newest result)."""
        if self.deferred and self.frame >= 1:
            self.ctx.exchange_wait(self._sync(self._last_counters), stream)

    def read(self, frame: int | None = None) -> dict:
        """Host view of the gathered arrays of the current frame (or of `frame`, while its set has not been reused), after a sync."""
        k = (self.frame if frame is None else frame) % self.nsets
        out = {}
        for name, dt in (("cmd", np.uint32), ("ptr", np.uint64), ("tag", np.uint32), ("counters", np.uint8)):
            buf = np.empty(self.sizes[name], np.uint8)
            self.ctx.memcpy_d2h(buf, self.local[k][name])
            out[name] = buf
        self.ctx.sync()
        cnt = out["counters"].reshape(self.world, self.counters_bytes)[:, 64:].copy().view(np.uint64)
        return dict(cmd=out["cmd"].view(np.uint32).reshape(-1, 5), ptr=out["ptr"].view(np.uint64).reshape(-1, 4),
                    tag=out["tag"].view(np.uint32).reshape(-1, 2), counts=cnt,
                    counters_raw=out["counters"].reshape(self.world, self.counters_bytes),
                    status=out["counters"].reshape(self.world, self.counters_bytes)[:, :4].copy().view(np.uint32)[:, 0])

    # -- consuming any rank's result on this GPU ----------------------------------------------------------------------
    def consume_params(self, r: int, inst_override: int | None = None):
        """cadr_cull_params describing rank r's result of the current frame AS SEEN FROM THIS GPU: commands, pointers and
        tags in slot r of the local gathered arrays, the counters rank r published, rank r's region table, and its
        instance indices + geometry / matrix lists through the peer mappings (addressDelta translates the addresses the
        records hold).  What a renderer on this GPU binds to draw rank r's part of the scene."""
        if not self.consumable:
            raise RuntimeError("PeerExchange was created without regions / inst_out / arena")
        p = self._capi.CullParams()
        st = self.local[self.frame % self.nsets]
        slot = r * self.cmd_cap
        p.numStateSets = self.num_ranges
        p.cmdOut, p.ptrOut, p.tagOut = st["cmd"] + 20 * slot, st["ptr"] + 32 * slot, st["tag"] + 8 * slot
        p.counters = st["counters"] + r * self.counters_bytes
        p.stateSetRegions = self.regions_dev[r]
        p.instOut = self.inst_of(r) if inst_override is None else inst_override
        p.addressDelta = self.peer_delta[r]
        return p

    def consume(self, r: int, stream: int = 0, pulled: bool = False) -> tuple[int, int]:
        """Walk EVERY range of rank r's result on this GPU the way the reference's vertex shader would
        (cadr_b200_consume_check_culled; shader.vert:99-123) -> (digest, fetches), summed over the ranges.
        pulled=True reads the instance indices from the local copies made by pull_instances instead of the peer mapping."""
        p = self.consume_params(r, self.gathered_inst + 4 * r * self.inst_cap if pulled else None)
        reg = self.peer_regions[r]
        live = [s for s in range(reg.shape[0]) if reg[s, 1]]
        for s in live:
            self.ctx.consume_check_culled(p, s, int(reg[s, 1]), self.digests + 16 * s, stream)
        out = np.zeros(2 * self.num_ranges, np.uint64)
        self.ctx.memcpy_d2h(out, self.digests, stream=stream)
        self.ctx.sync(stream)
        d = out.reshape(-1, 2)[live] if live else np.zeros((0, 2), np.uint64)
        return int(d[:, 0].sum(dtype=np.uint64)), int(d[:, 1].sum(dtype=np.uint64))

    # -- second stage (optional): the survivors' instance indices on the renderer GPU -------------------------------
    def enable_pull(self, inst_capacity: int) -> None:
        """Allocate this GPU's gathered instance-index buffer [world][capacity] (capacity = the largest rank's)."""
        caps = [None] * self.world
        dist.all_gather_object(caps, int(inst_capacity), group=self.group)
        self.inst_cap = (max(caps) + 3) & ~3
        self.gathered_inst = self.ctx.arena_alloc(max(self.world * self.inst_cap * 4, 256))

    def pull_instances(self, stream: int = 0, include_local: bool = False) -> None:
        """After end_frame: copy every peer's compacted instance-index runs of the current frame into gathered_inst
        (cadr_b200_exchange_pull_instances).  Rank r's commands then index gathered_inst + r * inst_cap * 4."""
        p = self._capi.ExchangePull()
        p.world, p.rank, p.numRanges, p.countersBytes = self.world, self.rank, self.num_ranges, self.counters_bytes
        p.gatheredCounters = self.local[self.frame % self.nsets]["counters"]
        p.gatheredInst, p.instCapacity, p.includeLocal = self.gathered_inst, self.inst_cap, int(include_local)
        for r in range(self.world):
            p.regions[r], p.peerInst[r] = self.regions_dev[r], self.inst_of(r)
        self.ctx.exchange_pull_instances(p, stream)

    # -- cross-check of the fused exchange over NCCL -------------------------------------------------------------------
    def verify(self, device, regions: list | None = None) -> dict:
        """What the peer stores of the cull kernels left in THIS rank's gathered arrays against what every rank holds
        for itself: each rank sends its own slot (commands, pointers, tags of every range in use, its counters) through
        an NCCL all-gather, and the received copies must equal, byte for byte, the slots the fused exchange filled here.
        -> dict(ok, commands, bytes, problems); identical on all ranks only if every rank's view is right, so callers
        all-reduce `ok`.  `regions`: every rank's region table [S,4], for an exchange created without them."""
        g = self.read()
        cap, cb = self.cmd_cap, self.counters_bytes
        me = slice(self.rank * cap, (self.rank + 1) * cap)
        payload = np.concatenate([g["cmd"][me].reshape(-1).view(np.uint8), g["ptr"][me].reshape(-1).view(np.uint8),
                                  g["tag"][me].reshape(-1).view(np.uint8), g["counters_raw"][self.rank].reshape(-1)])
        mine = torch.from_numpy(np.ascontiguousarray(payload)).to(device)
        everyone = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        got = everyone.cpu().numpy().reshape(self.world, -1)
        problems, commands, nbytes = [], 0, 0
        for r in range(self.world):
            o = got[r]
            cmd = o[:cap * 20].view(np.uint32).reshape(-1, 5)
            ptr = o[cap * 20:cap * 52].view(np.uint64).reshape(-1, 4)
            tag = o[cap * 52:cap * 60].view(np.uint32).reshape(-1, 2)
            ctr = o[cap * 60:cap * 60 + cb]
            if not np.array_equal(ctr, g["counters_raw"][r]):
                problems.append(f"counters of rank {r} differ from what it holds itself")
                continue
            counts = ctr[64:].view(np.uint64)
            reg = self.peer_regions[r] if self.consumable else (regions[r] if regions is not None else None)
            for s_ in range(len(counts)):
                c = int(counts[s_] & np.uint64(0xFFFFFFFF))
                if c == 0:
                    continue
                if reg is None:
                    raise RuntimeError("PeerExchange.verify needs the region tables (create it with regions=...)")
                b = int(reg[s_, 0])
                if b + c > cap or c > int(reg[s_, 1]):
                    problems.append(f"rank {r} range {s_}: {c} commands exceed the region")
                    continue
                lo, hi = r * cap + b, r * cap + b + c
                if not (np.array_equal(cmd[b:b + c], g["cmd"][lo:hi]) and np.array_equal(ptr[b:b + c], g["ptr"][lo:hi])
                        and np.array_equal(tag[b:b + c], g["tag"][lo:hi])):
                    problems.append(f"rank {r} range {s_}: records differ from the owner's")
                commands += c
                nbytes += c * 60
        return dict(ok=not problems, commands=commands, bytes=nbytes + self.world * cb, problems=problems[:8])

    def directory(self) -> list[dict]:
        """Host view after a sync: the draws a renderer issues for the whole scene, StateSet by StateSet - one
        indirect-count draw per (StateSet, rank that holds a piece of it): where the commands start in the gathered
        command buffer, how many there are (the count buffer entry), whose instance indices and matrices they refer to."""
        g = self.read()
        out = []
        for s in range(self.num_ranges):
            for r in range(self.world):
                c = int(g["counts"][r][s] & np.uint64(0xFFFFFFFF))
                if c:
                    reg = self.peer_regions[r] if self.consumable else (regions[r] if regions is not None else None)
                    out.append(dict(state_set=s, rank=r, count=c, instances=int(g["counts"][r][s] >> np.uint64(32)),
                                    first_command=r * self.cmd_cap + (int(reg[s, 0]) if reg is not None else 0)))
        return out

    def close(self) -> None:
        self.ctx.sync()
        dist.barrier(group=self.group)
        if self.consumable:
            for a in self.regions_dev:
                self.ctx.arena_free(a)
            self.ctx.arena_free(self.digests)
            self.regions_dev = []
        if getattr(self, "gathered_inst", 0):
            self.ctx.arena_free(self.gathered_inst)
            self.gathered_inst = 0
        for a in self._imported:
            self.ctx.ipc_close(a)
        self._imported = []
        for st in self.local:
            for a in st.values():
                self.ctx.arena_free(a)
        self.ctx.arena_free(self.flags)
        self.local = []
