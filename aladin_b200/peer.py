"""Peer windows over NVLink: one IPC-exportable device buffer per rank, mapped into every rank of the process group,
with copy-engine transfers and stream-ordered sequence flags (csrc/peer.cu).  Used by retrieval.AlignmentGallery to
replicate the packed caption rows of a phase on all ranks WHILE the persistent scoring kernel of the previous phase owns
the SMs -- an NCCL all-gather cannot start next to that kernel and ends up serialised between two scoring launches
(profiles/r02_e2e_timeline.md).  No counterpart in the reference (alad/evaluation.py is single-process)."""
import ctypes as C

import torch

from . import _cabi

WAIT_TIMEOUT_MS = 5000


class _Raw:
    """__cuda_array_interface__ holder: lets torch view memory this package allocated."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerWindow:
    """`nbytes` of zeroed device memory on every rank of `group`; ``ptrs[q]`` is rank q's buffer as seen from here.
    Creation and ``close`` are collective."""

    def __init__(self, nbytes, group):
        import torch.distributed as dist
        lib = _cabi.lib()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _cabi.MAX_PEERS:
            raise _cabi.AladError(f"peer windows support at most {_cabi.MAX_PEERS} ranks")
        self.nbytes = int(nbytes)
        self.device = torch.cuda.current_device()
        # every stage is followed by an exchange of its outcome, so that a failure on one rank raises on ALL ranks
        # instead of leaving the others inside a collective
        self.local, self.ptrs, self.bytes_view = None, [], None
        err, handle = None, (C.c_ubyte * _cabi.PEER_HANDLE_BYTES)()
        try:
            p = C.c_void_p()
            _cabi.check(lib.alad_peer_alloc(C.byref(p), self.nbytes), "alad_peer_alloc")
            self.local = p.value
            _cabi.check(lib.alad_peer_export(self.local, handle), "alad_peer_export")
        except _cabi.AladError as e:
            err = str(e)
        got = [None] * self.world
        dist.all_gather_object(got, (err, bytes(handle)), group=group)
        bad = [f"rank {q}: {e}" for q, (e, _) in enumerate(got) if e]
        if not bad:
            try:
                for q, (_, h) in enumerate(got):
                    if q == self.rank:
                        self.ptrs.append(self.local)
                        continue
                    buf = (C.c_ubyte * _cabi.PEER_HANDLE_BYTES).from_buffer_copy(h)
                    o = C.c_void_p()
                    _cabi.check(lib.alad_peer_open(buf, C.byref(o)), "alad_peer_open")
                    self.ptrs.append(o.value)
            except _cabi.AladError as e:
                err = str(e)
            got = [None] * self.world
            dist.all_gather_object(got, err, group=group)
            bad = [f"rank {q}: {e}" for q, e in enumerate(got) if e]
        if bad:
            for q, p in enumerate(self.ptrs):
                if q != self.rank:
                    lib.alad_peer_close(p)
            if self.local is not None:
                lib.alad_peer_free(self.local)
                self.local = None
            raise _cabi.AladError("peer window setup failed: " + "; ".join(bad))
        self.bytes_view = torch.as_tensor(_Raw(self.local, self.nbytes), device=torch.device("cuda", self.device))

    def view(self, offset, nbytes, dtype):
        """Typed 1-D view of the LOCAL buffer."""
        return self.bytes_view[offset:offset + nbytes].view(dtype)

    def close(self):
        import torch.distributed as dist
        if self.local is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)          # nobody copies into a buffer that is about to go away
        lib = _cabi.lib()
        for q, p in enumerate(self.ptrs):
            if q != self.rank:
                lib.alad_peer_close(p)
        self.bytes_view = None
        lib.alad_peer_free(self.local)
        self.local = None
        dist.barrier(group=self.group)


def copy(dst_ptr, src_ptr, nbytes, stream):
    _cabi.check(_cabi.lib().alad_peer_copy(dst_ptr, src_ptr, int(nbytes), stream.cuda_stream), "alad_peer_copy")


def signal(flag_ptrs, value, stream):
    """After everything enqueued on `stream`: store `value` into every listed flag slot (peer device pointers)."""
    n = len(flag_ptrs)
    if n == 0:
        return
    arr = (C.c_void_p * n)(*flag_ptrs)
    _cabi.check(_cabi.lib().alad_peer_signal(arr, n, int(value), stream.cuda_stream), "alad_peer_signal")


def wait(flags_ptr, n, value, skip, error_ptr, stream):
    """Block `stream` until the n local flag slots (except `skip`) have reached sequence number `value`."""
    _cabi.check(_cabi.lib().alad_peer_wait(flags_ptr, n, int(value), skip, WAIT_TIMEOUT_MS, error_ptr, stream.cuda_stream),
                "alad_peer_wait")
