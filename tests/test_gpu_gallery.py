"""Device-resident hand-off (aladin_b200.gallery): encode_data drop-in -> DeviceContainers ->
i2t / t2i / compute_recall.  Results must equal the host-container path bit for bit (same pack
kernel, same scoring kernel) and the reference's golden retrieval run."""
import contextlib
import io

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


class _FakeModel:
    """forward_emb contract of ALADModel (alad/alad_model.py:196-247 as consumed by
    alad/evaluation.py:114-130): sequence-major [S,B,d] token tensors + [B,d] global vectors."""

    def eval(self):
        return self

    def forward_emb(self, example_imgs, example_txts):
        img, img_len = example_imgs
        cap, cap_len = example_txts
        S_i, S_c = max(img_len), max(cap_len)
        img_emb = img[:, :S_i].permute(1, 0, 2).contiguous().cuda()
        cap_emb = cap[:, :S_c].permute(1, 0, 2).contiguous().cuda()
        return img[:, 0].cuda(), cap[:, 0].cuda(), img_emb, cap_emb, list(img_len), list(cap_len), None


class _Loader:
    def __init__(self, images, captions, img_lens, cap_lens, bs):
        self.dataset = range(images.shape[0])
        self.items = [((images[i:i + bs], img_lens[i:i + bs]), (captions[i:i + bs], cap_lens[i:i + bs]))
                      for i in range(0, images.shape[0], bs)]

    def __iter__(self):
        return iter(self.items)

    def __len__(self):
        return len(self.items)


def _containers(images, captions, img_lens, cap_lens, bs, precision):
    from aladin_b200 import gallery
    loader = _Loader(torch.from_numpy(images), torch.from_numpy(captions), img_lens, cap_lens, bs)
    return gallery.encode_data(_FakeModel(), loader, log_step=1000, logging=lambda *_: None, precision=precision)


@pytest.mark.parametrize("precision,bs", [("bf16", 7), ("fp32", 32)])
def test_device_gallery_equals_host_path(precision, bs):
    from aladin_b200 import evaluation, loss as L, synth
    from aladin_b200.recall_auxiliary import compute_recall
    images, captions, img_lens, cap_lens = synth.eval_containers(31, 64, 71, 96, max_regions=36, max_words=24)
    img_c, cap_c, il, cl = _containers(images, captions, img_lens, cap_lens, bs, precision)
    assert il == img_lens and cl == cap_lens and img_c.shape == (320, 71, 96)
    crit = L.AlignmentContrastiveLoss(aggregation="MrSw")
    crit.precision = precision

    def sim_fn(img, cap, img_len, cap_len):
        return crit(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)
    evaluation.clear_cache()
    ref_i = evaluation.i2t(ti, tc, img_lens, cap_lens, return_ranks=True, sim_function=sim_fn, cap_batches=5)
    ref_t = evaluation.t2i(ti, tc, img_lens, cap_lens, return_ranks=True, sim_function=sim_fn, im_batches=5)
    evaluation.clear_cache()
    got_i = evaluation.i2t(img_c, cap_c, il, cl, return_ranks=True, sim_function=sim_fn, cap_batches=5)
    got_t = evaluation.t2i(img_c, cap_c, il, cl, return_ranks=True, sim_function=sim_fn, im_batches=5)
    assert got_i[0] == ref_i[0] and got_t[0] == ref_t[0]
    for a, b in zip(got_i[1] + got_t[1], ref_i[1] + ref_t[1]):
        np.testing.assert_array_equal(a, b)
    # slot-0 access feeds the matching-head recall exactly like img_embs[:, 0, :] (alad/test.py:267)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        r_dev = compute_recall(img_c[:, 0, :], cap_c[:, 0, :])
        r_host = compute_recall(ti[:, 0, :], tc[:, 0, :])
    assert r_dev == r_host
    evaluation.clear_cache()
    g_dev = evaluation.i2t(img_c, cap_c, il, cl, sim_function=None)
    g_host = evaluation.i2t(ti, tc, img_lens, cap_lens, sim_function=None)
    assert g_dev == g_host
    with pytest.raises(TypeError):
        img_c[3]


def test_device_gallery_reproduces_reference_golden():
    from aladin_b200 import evaluation, loss as L
    g = load_golden("retrieval")
    images = np.repeat(g["images"], 5, axis=0)
    pad = np.zeros((images.shape[0], 71 - images.shape[1], images.shape[2]), np.float32)
    images71 = np.concatenate([images, pad], axis=1)                 # the reference zero-pads to 71 slots
    captions71 = np.concatenate([g["captions"], pad], axis=1)
    img_c, cap_c, il, cl = _containers(images71, captions71, g["img_lens"].tolist(), g["cap_lens"].tolist(), 16, "fp32")
    crit = L.AlignmentContrastiveLoss(aggregation="MrSw")
    crit.precision = "fp32"
    evaluation.clear_cache()
    m_i, (ranks_i, top1) = evaluation.i2t(img_c, cap_c, il, cl, return_ranks=True, sim_function=crit, cap_batches=5)
    m_t, (ranks_t, _) = evaluation.t2i(img_c, cap_c, il, cl, return_ranks=True, sim_function=crit, im_batches=5)
    # golden S was produced on [N,20,d] containers: every image there has masked slots or is full at 19 regions;
    # in the 71-slot container every image has masked slots (clamp at 0), which only matters for all-negative
    # columns -- none in this fixture, so ranks must be identical
    np.testing.assert_array_equal(ranks_i, g["ranks_i2t"])
    np.testing.assert_array_equal(ranks_t, g["ranks_t2i"])
    np.testing.assert_array_equal(top1, g["top1"])
    np.testing.assert_allclose(m_i[:5], g["m_i2t"][:5])
    np.testing.assert_allclose(m_t[:5], g["m_t2i"][:5])
