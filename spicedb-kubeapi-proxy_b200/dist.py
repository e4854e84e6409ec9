"""Multi-GPU plumbing for the check path: one process per GPU (torch.distributed).

The path shards by CHECK: every rank holds a replica of the snapshot (100 M tuples are
~2 GB, far below one GPU's HBM) and answers its own slice of the batch; there is no
data-path collective. The only communication is the
all-gather that returns the answers to the caller's order, plus the max-reduce of the
timings in bench.py. Works with backend "nccl" (GPU tensors) and "gloo" (CPU tensors;
used by the world_size-2 tests, where the evaluator is a stub).
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced slice of n items for this rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedChecker:
    """check_bulk over a process group: slice, answer locally, all-gather the answers.

    evaluate: callable(items ndarray) -> uint8 ndarray; on GPU ranks this is
    `engine.check_bulk`. Every rank must call check_bulk with the same items.
    """

    def __init__(self, evaluate: Callable[[np.ndarray], np.ndarray], group=None, device: str | None = None):
        self.evaluate = evaluate
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")

    def check_bulk(self, items: np.ndarray) -> np.ndarray:
        n = items.size
        lo, hi = shard_bounds(n, self.rank, self.world)
        mine = self.evaluate(items[lo:hi]) if hi > lo else np.empty(0, dtype=np.uint8)
        width = -(-n // self.world)  # every rank contributes a fixed-width block
        block = torch.zeros(width, dtype=torch.uint8)
        block[: hi - lo] = torch.from_numpy(np.ascontiguousarray(mine))
        block = block.to(self.device)
        gathered = [torch.empty_like(block) for _ in range(self.world)]
        dist.all_gather(gathered, block, group=self.group)
        out = np.empty(n, dtype=np.uint8)
        for r, t in enumerate(gathered):
            rlo, rhi = shard_bounds(n, r, self.world)
            out[rlo:rhi] = t[: rhi - rlo].cpu().numpy()
        return out


def max_over_ranks(value: float, group=None, device: str = "cpu") -> float:
    """Timings are reported as the max over ranks (bench.py contract)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


# ---------------------------------------------------------------------------------------------
# Object-hash sharded STORE (north-star's other multi-GPU mode).
#
# Every rank owns the relationships whose resource id % world == rank. A check starts on the
# owner of its resource; an edge to an object of another shard (or into a non-pure permission)
# is raised as a sub-query. Ranks run pass by pass; between passes the raised sub-queries are
# exchanged with one all-to-all (NCCL over NVLink on GPUs), until no rank raises any; values
# then travel back level by level with the reverse all-to-all and are folded into the jobs that
# raised them. Cross-shard traffic is 16 B per sub-query out, 1 B back.


class TorchTransport:
    """all-to-all of ragged byte arrays over torch.distributed (nccl: CUDA tensors, gloo: CPU)."""

    def __init__(self, group=None, device: str | None = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")

    def alltoall(self, send):
        """send[d]: 1-D uint8 array for rank d -> list recv[s] of uint8 arrays from every rank s."""
        counts = torch.tensor([int(a.size) for a in send], dtype=torch.int64, device=self.device)
        rcounts = torch.empty_like(counts)
        dist.all_to_all_single(rcounts, counts, group=self.group)
        rc = [int(x) for x in rcounts.cpu()]
        flat = np.concatenate([np.ascontiguousarray(a, dtype=np.uint8).ravel() for a in send]) if send else np.empty(0, np.uint8)
        tin = torch.from_numpy(flat.copy() if flat.size else np.zeros(0, np.uint8)).to(self.device)
        tout = torch.empty(sum(rc), dtype=torch.uint8, device=self.device)
        dist.all_to_all_single(tout, tin, output_split_sizes=rc, input_split_sizes=[int(a.size) for a in send],
                               group=self.group)
        out, host, off = [], tout.cpu().numpy(), 0
        for c in rc:
            out.append(host[off:off + c].copy())
            off += c
        return out

    def allreduce_sum(self, x: int) -> int:
        t = torch.tensor([int(x)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())


class LocalTransport:
    """The same interface for `world` virtual ranks living in ONE process (threads): used to
    test the sharded protocol on a single GPU. Build with LocalTransport.cluster(world)."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self._s = rank, world, shared

    @staticmethod
    def cluster(world):
        import threading

        shared = {"box": [None] * world, "bar": threading.Barrier(world), "sum": [0] * world}
        return [LocalTransport(r, world, shared) for r in range(world)]

    def alltoall(self, send):
        s = self._s
        s["box"][self.rank] = [np.ascontiguousarray(a, dtype=np.uint8).ravel().copy() for a in send]
        s["bar"].wait()
        out = [s["box"][src][self.rank] for src in range(self.world)]
        s["bar"].wait()
        return out

    def allreduce_sum(self, x):
        s = self._s
        s["sum"][self.rank] = int(x)
        s["bar"].wait()
        tot = sum(s["sum"])
        s["bar"].wait()
        return tot


class ShardedStoreChecker:
    """CheckBulkPermissions over an object-hash sharded store. Every rank calls check_bulk with the
    SAME items; every rank gets the full answer vector. `engine` must have been created with
    shard_rank=transport.rank, shard_count=transport.world and loaded with the full relationship
    stream (it keeps only what it owns)."""

    def __init__(self, engine, transport, check_dtype):
        self.e, self.t, self.dtype = engine, transport, np.dtype(check_dtype)
        self.stats = {"levels": 0, "subqueries_sent": 0, "bytes_sent": 0}

    def _route(self, subs):
        w = self.t.world
        dest = subs["res"] % np.uint32(w)
        groups = [np.flatnonzero(dest == d) for d in range(w)]  # stable partition, O(n * w), no sort
        order = np.concatenate(groups) if groups else np.empty(0, np.int64)
        return order, [subs[g].view(np.uint8) for g in groups]

    def check_bulk(self, items: np.ndarray) -> np.ndarray:
        items = np.ascontiguousarray(items, dtype=self.dtype)
        w, r = self.t.world, self.t.rank
        mine = np.nonzero(items["res"] % w == r)[0]
        queries = items[mine]
        levels = []
        while True:
            lv = len(levels)
            nsub = self.e.shard_pass(queries, lv)
            subs = self.e.shard_subqueries(lv, nsub)
            order, parts = self._route(subs)
            recv = self.t.alltoall(parts)
            levels.append({"nq": int(queries.size), "nsub": int(nsub), "order": order,
                           "recv_counts": [a.size // self.dtype.itemsize for a in recv]})
            self.stats["subqueries_sent"] += int(nsub)
            self.stats["bytes_sent"] += int(nsub) * self.dtype.itemsize
            queries = np.concatenate(recv).view(self.dtype) if recv else np.empty(0, self.dtype)
            if self.t.allreduce_sum(queries.size) == 0:
                break
            if lv > 60:
                raise RuntimeError("sharded check did not converge within the dispatch depth")
        self.stats["levels"] = max(self.stats["levels"], len(levels))
        out_next = None
        for lv in reversed(range(len(levels))):
            L = levels[lv]
            if lv == len(levels) - 1:
                child = np.empty(0, np.uint8)  # the deepest level raised nothing
            else:
                # values of the level lv+1 queries go back to the ranks that raised them
                rc = levels[lv]["recv_counts"]
                offs = np.concatenate([[0], np.cumsum(rc)])
                back = [out_next[offs[s]:offs[s + 1]] for s in range(w)]
                got = self.t.alltoall(back)
                flat = np.concatenate(got) if got else np.empty(0, np.uint8)
                child = np.empty(L["nsub"], np.uint8)
                child[L["order"]] = flat
                self.stats["bytes_sent"] += int(flat.size)
            out_next = self.e.shard_fold(lv, child, L["nq"])
        # everyone learns every answer: 4-byte index + 1-byte code per owned check
        payload = np.concatenate([mine.astype(np.uint32).view(np.uint8), out_next.astype(np.uint8)])
        got = self.t.alltoall([payload] * w)
        out = np.empty(items.size, np.uint8)
        for g in got:
            k = g.size // 5
            out[g[: 4 * k].view(np.uint32)] = g[4 * k:]
        return out


# ---------------------------------------------------------------------------------------------
# The same protocol with every buffer resident on the device (north-star's mode).
#
# Per level: ONE kernel pass over the level's queries (zg_shard_pass_dev), a counting-sort kernel that bucketises
# the raised sub-queries by owner (zg_shard_route_dev), an N-int count exchange and one all-to-all of the 16-byte
# sub-queries, device buffer to device buffer (NCCL over NVLink). Values (1 byte per sub-query) travel back with
# the reverse all-to-all and are OR-ed into the (query, leaf) that raised them by a kernel that reads them in
# routed order (zg_shard_fold_dev). Nothing is staged through the host; the host sees N counts per level.


class TorchDeviceTransport:
    """Ragged all-to-all of CUDA tensors over torch.distributed (backend nccl)."""

    def __init__(self, group=None, device=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device or torch.device("cuda", torch.cuda.current_device())

    def exchange_counts(self, counts):
        t = torch.tensor(counts, dtype=torch.int64, device=self.device)
        r = torch.empty_like(t)
        dist.all_to_all_single(r, t, group=self.group)
        return [int(x) for x in r.cpu()]

    def alltoall(self, send: torch.Tensor, send_counts, recv_counts, width: int) -> torch.Tensor:
        """send: uint8 tensor of sum(send_counts) * width bytes, destination-major."""
        out = torch.empty(sum(recv_counts) * width, dtype=torch.uint8, device=self.device)
        dist.all_to_all_single(out, send, output_split_sizes=[c * width for c in recv_counts],
                               input_split_sizes=[c * width for c in send_counts], group=self.group)
        torch.cuda.current_stream().synchronize()  # the engine's calls run on its own stream
        return out

    def allreduce_sum(self, x: int) -> int:
        t = torch.tensor([int(x)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())


class LocalDeviceTransport:
    """`world` virtual ranks in ONE process on ONE GPU (threads): tensors change hands by reference. Tests the
    device-resident protocol without NCCL. Build with LocalDeviceTransport.cluster(world)."""

    def __init__(self, rank, world, shared, device=None):
        self.rank, self.world, self._s = rank, world, shared
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()

    @staticmethod
    def cluster(world, device=None):
        """device=torch.device("cpu"): host tensors (the kernel emulator of tests/emu drives the same protocol)."""
        import threading

        shared = {"box": [None] * world, "bar": threading.Barrier(world), "sum": [0] * world}
        return [LocalDeviceTransport(r, world, shared, device) for r in range(world)]

    def exchange_counts(self, counts):
        s = self._s
        s["box"][self.rank] = list(counts)
        s["bar"].wait()
        out = [s["box"][src][self.rank] for src in range(self.world)]
        s["bar"].wait()
        return out

    def alltoall(self, send, send_counts, recv_counts, width):
        s = self._s
        offs = np.concatenate([[0], np.cumsum(send_counts)]) * width
        self._sync()
        s["box"][self.rank] = [send[int(offs[d]):int(offs[d + 1])] for d in range(self.world)]
        s["bar"].wait()
        parts = [s["box"][src][self.rank] for src in range(self.world)]
        out = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=self.device)
        self._sync()
        s["bar"].wait()
        return out

    def allreduce_sum(self, x):
        s = self._s
        s["sum"][self.rank] = int(x)
        s["bar"].wait()
        tot = sum(s["sum"])
        s["bar"].wait()
        return tot


class DeviceShardedChecker:
    """CheckBulkPermissions over an object-hash sharded store, device resident. Every rank brings ITS OWN batch
    (a CUDA uint8 tensor of n x 16 bytes) and gets the answers of that batch (CUDA uint8 tensor of n codes):
    checks are routed to the owner of their resource, evaluated level by level across the shards, and the
    answers routed back."""

    ITEM = 16

    def __init__(self, engine, transport):
        self.e, self.t = engine, transport
        self.stats = {"levels": 0, "subqueries_sent": 0, "bytes_sent": 0, "exchanges": 0}

    def _route(self, d_items: int, n: int, level: int):
        """-> (routed uint8 tensor, src int32 tensor, send_counts, recv_counts)"""
        dev, w = self.t.device, self.t.world
        routed = torch.empty(max(n, 1) * self.ITEM, dtype=torch.uint8, device=dev)
        src = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        counts = self.e.shard_route_dev(d_items, n, level, w, routed.data_ptr(), src.data_ptr())
        return routed[: n * self.ITEM], src[:n], counts, self.t.exchange_counts(counts)

    def check_bulk(self, d_items: torch.Tensor, n: int) -> torch.Tensor:
        dev, t = self.t.device, self.t
        # level "-1": this rank's own batch goes to the owners of its resources
        routed, src0, sc, rc = self._route(d_items.data_ptr(), n, -1)
        queries = t.alltoall(routed, sc, rc, self.ITEM)
        self.stats["exchanges"] += 1
        levels = [{"src": src0, "send": sc, "recv": rc}]  # routing that produced level lv's queries
        lv = 0
        while True:
            nq = queries.numel() // self.ITEM
            nsub = self.e.shard_pass_dev(queries.data_ptr() if nq else 0, nq, lv)
            routed, src, sc, rc = self._route(0, nsub, lv)
            nxt = t.alltoall(routed, sc, rc, self.ITEM)
            self.stats["exchanges"] += 1
            self.stats["subqueries_sent"] += int(nsub)
            self.stats["bytes_sent"] += int(nsub) * self.ITEM
            levels.append({"src": src, "send": sc, "recv": rc, "nq": nq, "nsub": int(nsub)})
            queries = nxt
            lv += 1
            if t.allreduce_sum(queries.numel()) == 0:
                break
            if lv > 60:
                raise RuntimeError("sharded check did not converge within the dispatch depth")
        self.stats["levels"] = max(self.stats["levels"], lv)
        # values flow back: level lv's outputs answer the sub-queries level lv-1 raised
        out_next = torch.empty(0, dtype=torch.uint8, device=dev)  # outputs of the (empty) deepest level
        for k in range(lv - 1, -1, -1):
            L = levels[k + 1]
            # out_next holds one byte per query that ARRIVED at level k+1, source-major: send them home
            back = t.alltoall(out_next, L["recv"], L["send"], 1)
            self.stats["bytes_sent"] += int(back.numel())
            out = torch.empty(max(L["nq"], 1), dtype=torch.uint8, device=dev)
            self.e.shard_fold_dev(k, back.data_ptr() if L["nsub"] else 0, L["src"].data_ptr() if L["nsub"] else 0, L["nsub"],
                                  out.data_ptr(), final_codes=(k == 0))
            out_next = out[: L["nq"]]
        # level 0's outputs answer the checks that were routed here: back to their callers, in the callers' order
        home = t.alltoall(out_next, levels[0]["recv"], levels[0]["send"], 1)
        answers = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        self.e.shard_unroute_dev(levels[0]["src"].data_ptr() if n else 0, home.data_ptr() if n else 0, n, answers.data_ptr())
        return answers[:n]
