"""Device-resident hand-off between the backbone and the retrieval path.

The reference's ``encode_data`` (alad/evaluation.py:80-155) copies every batch of backbone
outputs to the host into two zero-padded ``[N, 71, d]`` fp32 tensors (5.4 GB each at COCO-5k,
d = 768; images stored 5x), and ``i2t`` / ``t2i`` then push slices of them back to the GPU per
query.  Here the batches never leave the device: ``GalleryWriter.add_batch`` normalises, casts
and packs the scored tokens straight into the layout ``alad_mrsw_scores_fwd`` reads (the same
``alad_pack_tokens`` kernel the host path uses, so the scores are bit-identical), and keeps the
slot-0 global vectors for the matching head.  ``encode_data`` below is the drop-in with the
reference's signature; it returns two ``DeviceContainer`` objects that ``aladin_b200.evaluation.
i2t / t2i`` and ``aladin_b200.recall_auxiliary.compute_recall`` accept in place of the CPU tensors
(``container[:, 0, :]`` yields the global vectors, like ``img_embs[:, 0, :]`` at alad/test.py:267)."""
import time

import numpy as np
import torch

from . import _cabi, scoring
from .tiling import exclusive_cumsum, padded_rows, round_up, valid_counts

CONTAINER_SLOTS = 71            # max_img_len = max_cap_len = 71 (alad/evaluation.py:98-99)


class DeviceContainer:
    """Stands in for one ``[N, 71, d]`` container of ``encode_data``.

    kind 'images': packed rows hold the scored regions (slots 1 .. len-1) of the DISTINCT images,
    i.e. of items 0, 5, 10, ... (every image is stored 5x by the loader: dataset.py:117-119; i2t
    reads row 5i, t2i rows 0::5 -- evaluation.py:178,252).  kind 'captions': the scored words
    (slots 1 .. len-3) of every caption.  ``global_vecs`` [N, d] fp32 is slot 0 of every item."""

    def __init__(self, kind, n_items, d, precision, packed, counts, lengths, global_vecs, group):
        self.kind, self.n_items, self.d, self.precision = kind, n_items, d, precision
        self.packed = packed                   # scoring.Packed over the scored items
        self.counts = counts                   # int32 [n_scored_items] valid scored tokens
        self.lengths = lengths                 # raw python lengths of all N items
        self.global_vecs = global_vecs
        self.group = group                     # 5 for images (one scored item per 5 rows), 1 for captions

    # ---- just enough of the tensor protocol for the reference's call sites
    @property
    def shape(self):
        return (self.n_items, CONTAINER_SLOTS, self.d)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def is_cuda(self):
        return True

    @property
    def device(self):
        return self.packed.data.device

    def __len__(self):
        return self.n_items

    def __getitem__(self, key):
        """Only the slot-0 access of the call sites is supported: ``c[:, 0, :]`` / ``c[:, 0]``."""
        if isinstance(key, tuple) and len(key) >= 2 and key[1] == 0 and key[0] == slice(None) and \
                all(k == slice(None) for k in key[2:]):
            return self.global_vecs
        raise TypeError("DeviceContainer keeps only packed scored tokens and the slot-0 global vectors on the "
                        "device; index it as container[:, 0, :] or pass it whole to aladin_b200.evaluation.i2t / t2i")

    def rows_of(self, lo, hi):
        """(Packed view, counts) of scored items [lo, hi) -- a contiguous row range."""
        off = self.packed.row_off
        r0 = int(off[lo]) if lo < len(off) else self.packed.n_rows
        r1 = int(off[hi]) if hi < len(off) else self.packed.n_rows
        p = self.packed
        item = p.row_item[r0:] if p.row_item is not None else None
        return scoring.Packed(p.data[r0:r1] if r1 > r0 else p.data[:0], r1 - r0, p.Kp, p.counts[lo:hi],
                              off[lo:hi] - r0, item, p.mode), self.counts[lo:hi]


class GalleryWriter:
    """Accumulates backbone output batches on the device in the packed scoring layout."""

    def __init__(self, n_items, precision=None, device=None):
        if not torch.cuda.is_available():
            raise _cabi.AladError("aladin_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.N = int(n_items)
        self.precision = precision or scoring.get_precision()
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.pos = 0
        self.img_chunks, self.cap_chunks = [], []          # (Packed, counts) per batch
        self.img_lengths, self.cap_lengths = [], []
        self.img_glob = self.cap_glob = None
        self.d = None

    def add_batch(self, img_emb, cap_emb, img_cls, cap_cls, img_length, cap_length):
        """One ``model.forward_emb`` batch (alad/evaluation.py:114-130): img_emb [S_i,B,d], cap_emb
        [S_c,B,d] (sequence-major, as the backbone returns them), img_cls / cap_cls [B,d] global
        vectors, python length lists.  Tokens beyond slot 70 are dropped like the 71-slot container does."""
        split = scoring.PRECISION_CODE[self.precision]      # 0 bf16, 1 split-precision fp32, 2 tf32
        B = img_emb.shape[1]
        d = img_emb.shape[2]
        if self.d is None:
            self.d = d
            self.img_glob = torch.zeros((self.N, d), dtype=torch.float32, device=self.device)
            self.cap_glob = torch.zeros((self.N, d), dtype=torch.float32, device=self.device)
        if self.pos + B > self.N:
            raise ValueError("more items than the gallery was sized for")
        ids = np.arange(self.pos, self.pos + B)
        # ---- captions: scored words are slots 1 .. len-3 of the 71-slot container (W = 68)
        cap = scoring._require_cuda(cap_emb.detach(), "cap_emb").permute(1, 0, 2)[:, :CONTAINER_SLOTS]
        nw = valid_counts(cap_length, 3, CONTAINER_SLOTS - 3)
        nw = np.minimum(nw, max(cap.shape[1] - 1, 0)).astype(np.int32)     # slots the backbone did not produce are zero rows:
        # they would be scored as zero vectors by the reference; lengths never exceed the produced extent in practice
        self.cap_chunks.append((scoring.pack_tokens(cap, nw, slot0=1, mode=scoring.WORD_MODE[split]), nw))
        # ---- images: only the distinct ones (items 0, 5, 10, ...), scored regions are slots 1 .. len-1 (R = 70)
        sel = np.nonzero(ids % 5 == 0)[0]
        if sel.size:
            img = scoring._require_cuda(img_emb.detach(), "img_emb").permute(1, 0, 2)[:, :CONTAINER_SLOTS]
            first, step = int(sel[0]), 5
            img = img[first::step]
            nr = valid_counts([img_length[i] for i in sel], 1, CONTAINER_SLOTS - 1)
            nr = np.minimum(nr, max(img.shape[1] - 1, 0)).astype(np.int32)
            self.img_chunks.append((scoring.pack_tokens(img, nr, slot0=1, mode=scoring.REGION_MODE[split]), nr))
        self.img_glob[self.pos:self.pos + B] = img_cls.detach().to(self.device, torch.float32)
        self.cap_glob[self.pos:self.pos + B] = cap_cls.detach().to(self.device, torch.float32)
        self.img_lengths.extend(int(x) for x in img_length)
        self.cap_lengths.extend(int(x) for x in cap_length)
        self.pos += B

    def _concat(self, chunks, want_row_item):
        counts = np.concatenate([c for _, c in chunks]).astype(np.int32) if chunks else np.zeros(0, np.int32)
        row_off, n_rows = exclusive_cumsum(counts)
        Kp = chunks[0][0].Kp if chunks else round_up(self.d * scoring.K_FACTOR[scoring.PRECISION_CODE[self.precision]], _cabi.TILE_K)
        data = torch.empty((max(n_rows, 1), Kp), dtype=torch.bfloat16, device=self.device)
        r = 0
        for p, _ in chunks:
            if p.n_rows:
                data[r:r + p.n_rows].copy_(p.data[:p.n_rows])
            r += p.n_rows
        row_item = None
        if want_row_item:
            row_item = torch.full((max(padded_rows(n_rows), 2 * _cabi.TILE_M),), -1, dtype=torch.int32, device=self.device)
            if n_rows:
                row_item[:n_rows] = torch.repeat_interleave(
                    torch.arange(len(counts), dtype=torch.int32, device=self.device),
                    torch.from_numpy(counts.astype(np.int64)).to(self.device))
        mode = chunks[0][0].mode if chunks else 0
        return scoring.Packed(data, n_rows, Kp, counts, row_off, row_item, mode), counts

    def finalize(self):
        """-> (image DeviceContainer, caption DeviceContainer, img_lengths, cap_lengths)."""
        if self.pos != self.N:
            raise ValueError(f"gallery sized for {self.N} items, {self.pos} were added")
        img_p, nr = self._concat(self.img_chunks, False)
        cap_p, nw = self._concat(self.cap_chunks, True)
        self.img_chunks, self.cap_chunks = [], []
        imgs = DeviceContainer("images", self.N, self.d, self.precision, img_p, nr, self.img_lengths, self.img_glob, 5)
        caps = DeviceContainer("captions", self.N, self.d, self.precision, cap_p, nw, self.cap_lengths, self.cap_glob, 1)
        return imgs, caps, self.img_lengths, self.cap_lengths


def encode_data(model, data_loader, log_step=10, logging=print, precision=None):
    """Drop-in for ``alad.evaluation.encode_data`` (alad/evaluation.py:80-155): same arguments, same
    4-tuple, but the two containers are ``DeviceContainer`` objects that never leave the GPU."""
    model.eval()
    writer = GalleryWriter(len(data_loader.dataset), precision=precision)
    end = time.time()
    for i, batch_data in enumerate(data_loader):
        example_imgs, example_txts = batch_data
        with torch.no_grad():
            img_cross_attention, cap_cross_attention, img_emb, cap_emb, img_length, cap_length, _ = \
                model.forward_emb(example_imgs, example_txts)
            writer.add_batch(img_emb, cap_emb, img_cross_attention, cap_cross_attention, img_length, cap_length)
        if i % log_step == 0:
            logging('Test: [{0}/{1}]\tTime {2:.3f}'.format(i, len(data_loader), time.time() - end))
        end = time.time()
        del batch_data
    return writer.finalize()
