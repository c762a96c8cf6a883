"""CPU-only: the claim / completion protocol of the cross-GPU work pool (aladin_b200/steal.py) with the ranks as threads
of one process.  CUDA events are replaced by timers, launches by a log: every chunk of every rank must be executed exactly
once, fast ranks must take chunks of slow ones, nobody may touch a block before its owner has zeroed it, and every rank
must come back -- over several epochs, with ranks entering an epoch late."""
import threading
import time

import numpy as np
import pytest


class _Timer:
    """Stands in for a CUDA event: completes `delay` seconds after `start` (the end of the rank's simulated queue)."""

    def __init__(self, t_done):
        self.t_done = t_done

    def query(self):
        return time.perf_counter() >= self.t_done

    def synchronize(self):
        dt = self.t_done - time.perf_counter()
        if dt > 0:
            time.sleep(dt)


class _Pool:
    def __init__(self, cnt, rank, world):
        self.cnt, self.rank, self.world, self.epoch = cnt, rank, world, 0
        self.stats = {"own": 0, "stolen": 0}


@pytest.mark.parametrize("world,chunks", [(2, 4), (4, 8), (8, 8)])
def test_every_chunk_runs_once_and_slow_ranks_get_help(world, chunks):
    from aladin_b200 import steal
    words = np.zeros(4 * world + 8, dtype=np.int64)
    speed = [1.0 + 0.6 * (r == 1) + 0.3 * (r == world - 1) for r in range(world)]      # rank 1 is 60 % slower
    log = [[] for _ in range(world)]                # per epoch: list of (executor, owner, chunk, time)
    zero_at = {}
    errors = []
    lock = threading.Lock()
    epochs = 3

    def rank_main(r):
        try:
            pool = _Pool(steal.Counters(None, words=words, world=world, rank=r), r, world)
            for e in range(epochs):
                if r == world - 1 and e == 1:
                    time.sleep(0.004)               # enters the epoch late: its chunks must wait for it
                queue_end = [time.perf_counter()]

                def enqueue(ms):
                    queue_end[0] = max(queue_end[0], time.perf_counter()) + ms * 1e-3 * speed[r]
                    return _Timer(queue_end[0])

                def zero_event():
                    ev = enqueue(0.3)
                    with lock:
                        zero_at[(e, r)] = ev.t_done
                    return ev

                def launch_chunk(owner, c):
                    with lock:
                        log[e % len(log)].append((e, r, owner, c, time.perf_counter()))
                    last[0] = enqueue(1.0)

                last = [None]
                steal.run(pool, chunks, lambda: enqueue(6.0), launch_chunk, 0, zero_event, make_event=lambda: last[0])
                barrier.wait()                      # the ranking exchange: all ranks meet before the next epoch
        except Exception as ex:                     # pragma: no cover
            errors.append(repr(ex))
            barrier.abort()

    barrier = threading.Barrier(world)
    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads), "a rank did not come back"
    entries = [x for l in log for x in l]
    for e in range(epochs):
        got = sorted((o, c) for (ee, _, o, c, _) in entries if ee == e)
        assert got == sorted((o, c) for o in range(world) for c in range(chunks)), f"epoch {e}: chunks lost or run twice"
        for (ee, ex, o, c, t) in entries:
            if ee == e and ex != o:
                assert t >= zero_at[(e, o)] - 1e-4, "a chunk was taken before its owner had zeroed the block"
    helped = [(ee, ex, o) for (ee, ex, o, c, _) in entries if ex != o]
    assert any(o == 1 for (_, _, o) in helped), "the slow rank got no help"
