"""Multi-GPU search: one process per GPU, `torch.distributed` for the plumbing.

The reference parallelises search_5lut / search_7lut over MPI ranks with a master/worker protocol
(lut.c:137-159, 212-238, 329-360, 463-482, 664-740; sboxgates.c:619-642): contiguous slices of the
combination space, first finder wins.  Here every rank runs the same host program with the same
RNG state, takes an interleaved share of the work items, and the ranks agree on the answer with
  * search_5lut: one all-reduce(MIN) of the 64-bit key (rank of combination, ordering, position);
  * search_7lut: one all-gather of the per-rank hit lists (phase 1, lut.c:329-349), then one
    all-reduce(MIN) of the key (list index, ordering, outer position, middle position).
Because the key orders candidates exactly as the reference's single-rank loop visits them, the
result is the reference's size == 1 result for any number of GPUs.

The engine is any object with the `LutEngine` part-methods; the CPU tests drive this module over
`gloo` with an oracle-backed stand-in engine, the product uses `LutEngine` (CUDA) over `nccl`.
"""
import math
import time

import numpy as np
import torch
import torch.distributed as dist

from .lut import (SBG_KEY_NONE, SBG_LIST_CAP, result5_to_ret, result7_to_ret, shuffled_order,
                  shuffled_orders7)

_I64_MAX = (1 << 63) - 1


def _key_to_i64(key):
    # Keys use < 2^63 except the "none" sentinel, which maps to int64 max (still the maximum).
    return _I64_MAX if key == SBG_KEY_NONE else int(key)


def _i64_to_key(v):
    return SBG_KEY_NONE if v == _I64_MAX else int(v)


class _DeviceArray:
    """A raw device pointer as something torch can wrap without a copy."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (ptr, False),
                                         "version": 2}


class DistributedLutSearch:
    """Sharding pays only when a search is large: below the thresholds every rank simply runs the
    whole (sub)search itself -- same inputs, same deterministic result, no collective at all."""

    def __init__(self, engine, group=None, device=None, shard_min_tuples5=5e7,
                 shard_min_tuples7=2e8, shard_min_list=8192):
        self.engine = engine
        self.shard_min_tuples5 = shard_min_tuples5
        self.shard_min_tuples7 = shard_min_tuples7
        self.shard_min_list = shard_min_list
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" \
                else torch.device("cpu")
        self.device = device
        self.collectives = 0
        self.collective_ms = 0.0      # host time spent inside collectives (incl. waiting for peers)
        self.last_phase1_sharded = False

    # -- collectives ---------------------------------------------------------------------------
    def _allreduce_min_key(self, key):
        t0 = time.perf_counter()
        t = torch.tensor([_key_to_i64(key)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        self.collectives += 1
        out = _i64_to_key(int(t.item()))
        self.collective_ms += 1e3 * (time.perf_counter() - t0)
        return out

    def _allgather_lists(self, local):
        """local: sorted uint64 array (<= SBG_LIST_CAP).  Returns the concatenation over ranks.
        Host path (gloo / engines without a device-side list)."""
        cnt = torch.tensor([local.shape[0]], dtype=torch.int64, device=self.device)
        counts = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(counts, cnt, group=self.group)
        counts = [int(c.item()) for c in counts]
        width = max(counts)
        self.collectives += 1
        if width == 0:
            return np.zeros(0, dtype=np.uint64)
        buf = torch.zeros(width, dtype=torch.int64)
        buf[:local.shape[0]] = torch.from_numpy(local.view(np.int64))
        buf = buf.to(self.device)
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf, group=self.group)
        self.collectives += 1
        out = [p[:c].cpu().numpy().view(np.uint64) for p, c in zip(parts, counts)]
        return np.concatenate(out)

    def _allgather_merge_on_device(self, count):
        """NCCL path: every rank's ordered list goes straight from its engine's device buffer into
        one gathered device buffer (all_gather_into_tensor over NVLink) and is merged there
        (sbg_set_list7_device); no list ever visits the host.  Returns the merged length."""
        t0 = time.perf_counter()
        cnt = torch.tensor([count], dtype=torch.int64, device=self.device)
        counts = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(counts, cnt, group=self.group)
        counts = counts.tolist()          # `world` integers: grid sizes are chosen on the host
        self.collectives += 1
        width = max(max(counts), 1)
        mine = torch.zeros(width, dtype=torch.int64, device=self.device)
        if count > 0:
            ptr, n = self.engine.list7_device()
            mine[:count] = torch.as_tensor(_DeviceArray(ptr, n), device=self.device)[:count]
        gathered = torch.empty(self.world * width, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        self.collectives += 1
        # the merge kernel runs on the engine's stream, the collective on torch's: order them
        torch.cuda.current_stream().synchronize()
        self.collective_ms += 1e3 * (time.perf_counter() - t0)
        self.engine.set_list7_device(gathered.data_ptr(), width, counts)
        self._keepalive = gathered
        return min(sum(counts), SBG_LIST_CAP)

    # -- searches ------------------------------------------------------------------------------
    def search5_sharded(self, order):
        """The current problem's 5-LUT search with a given function order -> raw sbg_result."""
        n = self.engine.n
        if self.world == 1 or math.comb(n, 5) < self.shard_min_tuples5:
            return self.engine.finish5(self.engine.search5_part(0, 1, order), order)
        key = self.engine.search5_part(self.rank, self.world, order)
        key = self._allreduce_min_key(key)
        return self.engine.finish5(key, order)

    def search7_sharded(self, outer, middle):
        n = self.engine.n
        self.last_phase1_sharded = not (self.world == 1 or math.comb(n, 7) < self.shard_min_tuples7)
        if not self.last_phase1_sharded:
            # phase 1 replicated: every rank builds (and keeps on its device) the same full list
            count = self.engine.filter7_keep_local()
        elif self.device.type == "cuda" and hasattr(self.engine, "filter7_part_device"):
            count = self._allgather_merge_on_device(
                self.engine.filter7_part_device(self.rank, self.world))
        else:
            local = self.engine.filter7_part(self.rank, self.world)
            merged = self._allgather_lists(local)
            # Every rank installs the same merged list: the runs are merged and cut at
            # SBG_LIST_CAP entries (lut.c:316-318 at size == 1).
            self.engine.set_list7(merged)
            count = min(len(merged), SBG_LIST_CAP)
        if self.world == 1 or count < self.shard_min_list:
            key = self.engine.decomp7_part(0, 1, outer, middle)
        else:
            key = self.engine.decomp7_part(self.rank, self.world, outer, middle)
            key = self._allreduce_min_key(key)
        return self.engine.finish7(key, outer, middle)

    def search_5lut(self, tables, target, mask, inbits, rng):
        order = shuffled_order(rng)
        self.engine.load(tables, target, mask, inbits)
        return result5_to_ret(self.search5_sharded(order), rng)

    def search_7lut(self, tables, target, mask, inbits, rng):
        outer, middle = shuffled_orders7(rng)
        self.engine.load(tables, target, mask, inbits)
        return result7_to_ret(self.search7_sharded(outer, middle), rng)
