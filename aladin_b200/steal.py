"""Cross-GPU work pool for the sharded scoring pass (no counterpart in the reference, which runs
one process).

Every rank owns a block of gallery images and scores it against ALL captions.  Under the 1 kW power cap the B200s of one
box do not run at the same clock, and not at a constant one: an 8-way shard's step takes 40.4 .. 45.6 ms on the SAME GPU
from one step to the next (profiles/r02_n8_work_pool.md), so the step of the job -- the maximum over the ranks -- sits
1.5-2 ms above the mean and no static partition removes that.  Here the last part of every rank's work (a range of
256-row word units, cut into chunks) goes into a pool: a rank that has finished its own chunks claims chunks of the
rank with the most work left, scores them from ITS copy of that rank's packed regions, and adds the partial sums
straight into the owner's score block over NVLink (the block lives in an IPC peer window; the epilogue's
red.global.add.f32 works on peer memory).  Claims and completions are counters in a /dev/shm page shared by the ranks
(host atomics, csrc/cabi.cu); the launches stay ordinary host-side launches of alad_mrsw_scores_fwd with
``accumulate = 1``, two chunks deep, so nothing spins on the device.  A caption whose rows straddle a chunk boundary
gets its two partial sums from two launches (possibly two GPUs): an fp32 sum of two addends does not depend on their
order, so the scores stay bit-identical to the unsharded single launch."""
import mmap
import os
import time

import numpy as np
import torch

from . import _cabi, peer

EPOCH_SHIFT = 12            # counter word = epoch << 12 | count
_MASK = (1 << EPOCH_SHIFT) - 1


class Counters:
    """[world, 4] int64 words in a shared host page: claim, done, ready (epoch), spare."""

    def __init__(self, group, words=None, world=None, rank=None):
        self.lib = _cabi.lib()
        if words is not None:                    # tests: ranks as threads of one process on a caller-owned int64 array
            self.world, self.rank, self.words, self.base = world, rank, words, words.ctypes.data
            self.map = self.fd = None
            return
        import torch.distributed as dist
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        name = [None]
        if self.rank == 0:
            try:
                name[0] = f"/dev/shm/alad_b200_pool_{os.getpid()}_{int(time.time() * 1e6) & 0xffffffff}"
                with open(name[0], "wb") as f:
                    f.write(b"\0" * mmap.PAGESIZE)
            except OSError as e:
                name[0] = f"!{e}"
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if hasattr(dist, "get_global_rank") else 0, group=group)
        self.path = name[0]
        err = self.path[1:] if self.path.startswith("!") else None
        self.map = self.fd = None
        if err is None:
            try:
                self.fd = os.open(self.path, os.O_RDWR)
                self.map = mmap.mmap(self.fd, mmap.PAGESIZE)
            except OSError as e:                 # e.g. ranks in different IPC / mount namespaces
                err = str(e)
        errs = [None] * self.world
        dist.all_gather_object(errs, err, group=group)     # a failure anywhere raises everywhere (also the barrier before unlink)
        if any(errs):
            raise _cabi.AladError("work-pool counters unavailable: " + "; ".join(f"rank {q}: {e}" for q, e in enumerate(errs) if e))
        self.words = np.frombuffer(self.map, dtype=np.int64)
        self.base = self.words.ctypes.data
        if self.rank == 0:
            os.unlink(self.path)                # the mappings keep the page alive

    def _p(self, owner, slot):
        return self.base + 8 * (4 * owner + slot)

    def load(self, owner, slot):
        return int(self.lib.alad_host_atomic_load(self._p(owner, slot)))

    def store(self, owner, slot, v):
        self.lib.alad_host_atomic_store(self._p(owner, slot), int(v))

    def add(self, owner, slot, v):
        return int(self.lib.alad_host_atomic_add(self._p(owner, slot), int(v)))

    def cas(self, owner, slot, old, new):
        return bool(self.lib.alad_host_atomic_cas(self._p(owner, slot), int(old), int(new)))

    def close(self):
        self.words = None
        if self.map is None:
            return
        try:
            self.map.close()
        except BufferError:
            pass
        os.close(self.fd)


CLAIM, DONE, READY = 0, 1, 2


def claim_chunk(cnt, epoch, owner, n_chunks):
    """Next unclaimed chunk of `owner` in this epoch, or None (pool empty, or the owner has not opened the epoch)."""
    while True:
        v = cnt.load(owner, CLAIM)
        if (v >> EPOCH_SHIFT) != epoch or (v & _MASK) >= n_chunks:
            return None
        if cnt.cas(owner, CLAIM, v, v + 1):
            return v & _MASK


def remaining(cnt, epoch, owner, n_chunks):
    v = cnt.load(owner, CLAIM)
    if (v >> EPOCH_SHIFT) != epoch or cnt.load(owner, READY) != epoch:
        return 0
    return max(0, n_chunks - (v & _MASK))


class WorkPool:
    """Per process group: counters, the peer window holding every rank's score block, and the epoch."""

    def __init__(self, group, block_rows, Nc):
        import torch.distributed as dist
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.block_rows, self.Nc = int(block_rows), int(Nc)
        self.cnt = Counters(group)
        self.win = peer.PeerWindow(max(self.block_rows * self.Nc * 4, 1024), group)
        self.epoch = 0
        self.stats = {"own": 0, "stolen": 0}

    def fits(self, rows, Nc):
        return rows <= self.block_rows and Nc == self.Nc

    def block(self, rows):
        """This rank's score block [rows, Nc] inside its window."""
        return self.win.view(0, rows * self.Nc * 4, torch.float32).view(rows, self.Nc)

    def close(self):
        self.win.close()
        self.cnt.close()


def _cuda_event():
    ev = torch.cuda.Event()
    ev.record()
    return ev


def run(pool, n_chunks, launch_main, launch_chunk, n_own_rows, zero_event_fn, make_event=_cuda_event):
    """One scoring pass of this rank under the pool protocol.

    launch_main()                 enqueue the main part of this rank's own block (zeroing first)
    launch_chunk(owner, chunk)    enqueue tail chunk `chunk` of `owner`'s block onto the owner's window; returns nothing
    zero_event_fn()               -> CUDA event recorded after this rank's block was zeroed (its chunks may be stolen from then on)
    Returns the number of (own, stolen) chunks this rank executed."""
    cnt, rank, world = pool.cnt, pool.rank, pool.world
    pool.epoch += 1
    epoch = pool.epoch
    tag = epoch << EPOCH_SHIFT
    cnt.store(rank, DONE, tag)
    cnt.store(rank, CLAIM, tag)                 # opens the epoch: from here on others may claim (once READY says so)
    zero_ev = zero_event_fn()
    launch_main()
    ready = False
    inflight = []                               # (owner, event), oldest first
    own = stolen = 0
    exhausted = False
    while True:
        if not ready and zero_ev.query():
            cnt.store(rank, READY, epoch)
            ready = True
        while len(inflight) < 2 and not exhausted:
            c = claim_chunk(cnt, epoch, rank, n_chunks)
            owner = rank
            if c is None:
                # own pool empty: help the rank with the most unclaimed chunks
                best, left = None, 0
                for q in range(world):
                    if q != rank:
                        n = remaining(cnt, epoch, q, n_chunks)
                        if n > left:
                            best, left = q, n
                if best is not None:
                    c = claim_chunk(cnt, epoch, best, n_chunks)
                    owner = best
                if c is None:
                    exhausted = best is None    # nothing claimable anywhere right now
                    break
            launch_chunk(owner, c)
            inflight.append((owner, make_event()))
            if owner == rank:
                own += 1
            else:
                stolen += 1
        if not inflight:
            if exhausted or all(remaining(cnt, epoch, q, n_chunks) == 0 for q in range(world) if q != rank):
                # ranks that have not opened the epoch yet keep their chunks for themselves or later thieves
                break
            time.sleep(20e-6)
            continue
        if not ready:                           # the block is zeroed ~2 ms into the step: publish before blocking on a chunk
            zero_ev.synchronize()
            cnt.store(rank, READY, epoch)
            ready = True
        owner, ev = inflight.pop(0)
        ev.synchronize()
        cnt.add(owner, DONE, 1)
        exhausted = False
    if not ready:
        zero_ev.synchronize()
        cnt.store(rank, READY, epoch)
    # every chunk of this rank's block has to be in before its scores are read
    while (cnt.load(rank, DONE) & _MASK) < n_chunks:
        time.sleep(10e-6)
    pool.stats["own"] += own
    pool.stats["stolen"] += stolen
    return own, stolen
