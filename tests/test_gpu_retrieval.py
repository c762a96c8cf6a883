"""GPU parity tests for the retrieval drop-ins (aladin_b200.evaluation / recall_auxiliary):
Recall@K, medr, meanr, ranks, top1 and top50 identical to the reference's outputs."""
import numpy as np
import pytest
import torch

from conftest import assert_order_equal_up_to_ties, assert_ranks_equal_up_to_ties, load_golden
from oracle import alad_oracle as O

pytestmark = pytest.mark.gpu


def _containers(g):
    images = torch.from_numpy(np.repeat(g["images"], 5, axis=0))
    captions = torch.from_numpy(g["captions"])
    return images, captions, g["img_lens"].tolist(), g["cap_lens"].tolist()


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
def test_i2t_t2i_alignment_golden(precision):
    from aladin_b200 import evaluation as E, loss as L
    g = load_golden("retrieval")
    images, captions, il, cl = _containers(g)
    sim_matrix_class = L.AlignmentContrastiveLoss(aggregation="MrSw")
    sim_matrix_class.precision = precision

    def alignment_sim_fn(img, cap, img_len, cap_len):             # the closure alad/test.py:259-263 builds
        with torch.no_grad():
            return sim_matrix_class(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

    m, (ranks, top1) = E.i2t(images, captions, il, cl, return_ranks=True, sim_function=alignment_sim_fn, cap_batches=5)
    mi, (ranks_i, top50) = E.t2i(images, captions, il, cl, return_ranks=True, sim_function=alignment_sim_fn, im_batches=5)
    np.testing.assert_allclose(m[:3], g["m_i2t"][:3])             # Recall@1/5/10 identical
    np.testing.assert_allclose(mi[:3], g["m_t2i"][:3])
    if precision == "fp32":
        np.testing.assert_allclose(m, g["m_i2t"])
        np.testing.assert_allclose(mi, g["m_t2i"])
        np.testing.assert_array_equal(ranks, g["ranks_i2t"])
        np.testing.assert_array_equal(top1, g["top1"])
        np.testing.assert_array_equal(ranks_i, g["ranks_t2i"])
        assert_order_equal_up_to_ties(top50, g["top50"], g["S_full"].T, 1e-4, "t2i top50")
    assert ranks.dtype == np.float64 and top50.shape == (300, 50)


def test_i2t_t2i_global_vector_and_compute_recall_golden(capsys):
    import aladin_b200
    from aladin_b200 import evaluation as E, recall_auxiliary as R
    g = load_golden("retrieval")
    images, captions, il, cl = _containers(g)
    aladin_b200.set_precision("fp32")
    try:
        m, (ranks, top1) = E.i2t(images, captions, il, cl, return_ranks=True, sim_function=None)
        mi, (ranks_i, top50) = E.t2i(images, captions, il, cl, return_ranks=True, sim_function=None)
        Sg = images[0::5][:, 0, :].numpy().astype(np.float64) @ captions[:, 0, :].numpy().astype(np.float64).T
        np.testing.assert_allclose(m[:3], g["g_i2t"][:3])            # Recall@1/5/10 identical
        np.testing.assert_allclose(mi[:3], g["g_t2i"][:3])
        np.testing.assert_array_equal(top1, g["gtop1"])
        gt_i2t = [5 * i + int(np.argmax(Sg[i, 5 * i:5 * i + 5])) for i in range(Sg.shape[0])]
        assert_ranks_equal_up_to_ties(ranks, g["granks_i2t"], Sg, gt_i2t, 1e-4, "global i2t ranks")
        assert_ranks_equal_up_to_ties(ranks_i, g["granks_t2i"], Sg.T, np.arange(Sg.shape[1]) // 5, 1e-4, "global t2i ranks")
        assert_order_equal_up_to_ties(top50, g["gtop50"], Sg.T, 1e-4, "global t2i top50")
        rec = R.compute_recall(images[:, 0, :], captions[:, 0, :])
        np.testing.assert_allclose(rec, g["compute_recall"])               # R@K of both directions + rsum
        assert "Recall Image to text" in capsys.readouterr().out
    finally:
        aladin_b200.set_precision("bf16")


def test_arbitrary_callable_sim_function():
    """A user callable is honoured (called per query like evaluation.py:199-210)."""
    from aladin_b200 import evaluation as E
    g = load_golden("retrieval")
    images, captions, il, cl = _containers(g)
    calls = []

    def sim(img, cap, img_len, cap_len):
        calls.append(cap.shape[0])
        return torch.from_numpy(O.mrsw_scores(img.cpu().numpy(), cap.cpu().numpy(), img_len, cap_len))

    m, (ranks, top1) = E.i2t(images, captions, il, cl, return_ranks=True, sim_function=sim, cap_batches=5)
    np.testing.assert_array_equal(ranks, g["ranks_i2t"])
    assert len(calls) == 60 * 5 and set(calls) == {60}


def test_closure_that_post_processes_the_scores_is_not_bypassed():
    """a closure over the drop-in criterion that changes its scores (here: adds the matching term, the
    commented-out `+ d_matching` of alad/evaluation.py:206-210) must be CALLED, not replaced by the fused block; the
    plain closure of alad/test.py:259-263 and the tagged helper take the fused path."""
    from aladin_b200 import evaluation as E, loss as L
    g = load_golden("retrieval")
    images, captions, il, cl = _containers(g)
    crit = L.AlignmentContrastiveLoss(aggregation="MrSw")
    crit.precision = "fp32"
    n_calls = []

    def plain(img, cap, img_len, cap_len):
        n_calls.append(img.shape[0])
        return crit(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

    def with_matching(img, cap, img_len, cap_len):
        n_calls.append(img.shape[0])
        d = crit(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)
        return d - 3.0 * torch.mm(img[:, 0, :], cap[:, 0, :].t())

    E.clear_cache()
    m_plain, (r_plain, _) = E.i2t(images, captions, il, cl, return_ranks=True, sim_function=plain, cap_batches=5)
    assert len(n_calls) == 1 and E._cache["key"][-2] == "fused:MrSw"          # the probe call only
    np.testing.assert_array_equal(r_plain, g["ranks_i2t"])
    n_calls.clear()
    E.clear_cache()
    m_post, (r_post, _) = E.i2t(images, captions, il, cl, return_ranks=True, sim_function=with_matching, cap_batches=5)
    assert len(n_calls) == 1 + 60 * 5                                         # probe, then one call per query and chunk
    S = g["S_full"].astype(np.float64) - 3.0 * (g["images"][:, 0, :].astype(np.float64) @ g["captions"][:, 0, :].astype(np.float64).T)
    ri, _ = O.i2t_ranks(S.astype(np.float32))
    gt = np.array([5 * i + int(np.argmax(S[i, 5 * i:5 * i + 5])) for i in range(60)])
    assert_ranks_equal_up_to_ties(r_post, ri, S.astype(np.float32), gt, 1e-4, "post-processed closure: i2t ranks")
    assert not np.array_equal(r_post, r_plain)
    n_calls.clear()
    E.clear_cache()
    tagged = E.fused_sim_function(crit)
    _, (r_tag, _) = E.i2t(images, captions, il, cl, return_ranks=True, sim_function=tagged, cap_batches=5)
    np.testing.assert_array_equal(r_tag, g["ranks_i2t"])


def test_coco1k_shape_subset_vs_oracle():
    """Dense 34x50, d=1024 tokens (BASELINE per-pair shape) at 100 images x 500 captions."""
    from aladin_b200 import evaluation as E, loss as L, synth
    images, captions, il, cl = synth.eval_containers(33, 100, 53, 1024, max_regions=35, max_words=53, dense=True, alpha=0.04)
    scorer = L.AlignmentContrastiveLoss(aggregation="MrSw")
    scorer.precision = "fp32"
    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)
    m, (ranks, top1) = E.i2t(ti, tc, il, cl, return_ranks=True, sim_function=scorer)
    mi, (ranks_i, top50) = E.t2i(ti, tc, il, cl, return_ranks=True, sim_function=scorer)
    S = O.mrsw_scores(images[0::5], captions, il[0::5], cl, acc64=True)
    ri, t1 = O.i2t_ranks(S)
    rt, t50 = O.t2i_ranks(S)
    np.testing.assert_array_equal(ranks, ri)
    np.testing.assert_array_equal(ranks_i, rt)
    np.testing.assert_allclose(m[:5], O.recall_metrics(ri))
    np.testing.assert_allclose(mi[:5], O.recall_metrics(rt))
    assert 5.0 < m[0] < 100.0                                       # non-degenerate recalls


def test_pinned_host_gallery_chunked_upload_matches_device_resident():
    from aladin_b200 import retrieval, synth
    images, captions, il, cl = synth.eval_containers(34, 64, 71, 256, max_regions=35, max_words=40)
    ti = torch.from_numpy(images).pin_memory()
    tc = torch.from_numpy(captions).pin_memory()
    a = retrieval.AlignmentGallery(ti, tc, il, cl, n_images=64, img_start=0, img_step=5, precision="bf16",
                                   caption_chunk=50).scores()
    b = retrieval.AlignmentGallery(ti.cuda(), tc.cuda(), il, cl, n_images=64, img_start=0, img_step=5,
                                   precision="bf16").scores()
    torch.cuda.synchronize()
    # every upload chunk starts at its canonical row modulo the kernel's 256-row work unit, so a caption's words are cut
    # into the same <= 2 partial sums however the gallery arrives: bit-identical scores
    assert torch.equal(a, b)
    ref = O.mrsw_scores(images[0::5], captions, il[0::5], cl, acc64=True)
    assert np.abs(a.cpu().numpy() - ref).max() <= 1e-2


def test_pageable_host_gallery_staged_upload_matches_device_resident(monkeypatch):
    """The tensors the reference's encode_data returns are PAGEABLE: their rows reach the device through pinned staging
    buffers filled by several host threads (alad_h2d_2d_staged).  82 MB of captions = two 48 MB staging buffers per chunk,
    5-row pitch on the image side, odd thread count; the scores must equal the device-resident gallery's bit for bit."""
    from aladin_b200 import retrieval, synth
    monkeypatch.setenv("ALAD_H2D_THREADS", "3")
    images, captions, il, cl = synth.eval_containers(35, 800, 21, 256, max_regions=20, max_words=17)
    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)
    assert not ti.is_pinned() and not tc.is_pinned() and captions.nbytes > (48 << 20)
    a = retrieval.AlignmentGallery(ti, tc, il, cl, n_images=800, img_start=0, img_step=5, precision="bf16",
                                   caption_chunk=4000).scores()
    b = retrieval.AlignmentGallery(ti.cuda(), tc.cuda(), il, cl, n_images=800, img_start=0, img_step=5,
                                   precision="bf16").scores()
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    # and a direct check of the upload itself: every 5th image row, the first 9 slots
    dev = retrieval._upload_rows(ti, 0, 5, 800, 9)
    assert torch.equal(dev.cpu(), ti[0::5, :9])


@pytest.mark.parametrize("block,precision", [(64, "bf16"), (150, "fp32"), (1000, "bf16")])
def test_streaming_ranks_equal_the_dense_matrix(block, precision):
    """Block-by-block retrieval without the [Ni, Nc] matrix (retrieval.streaming_ranks): ranks, top-1 and top-50 must equal the
    ranking of the dense score matrix exactly -- ragged lengths, an image without regions, captions without words, blocks that
    do not divide the gallery."""
    from aladin_b200 import retrieval, synth
    Ni = 333
    images, captions, il, cl = synth.eval_containers(36, Ni, 40, 192, max_regions=36, max_words=38, alpha=0.3)
    il[5 * 7:5 * 7 + 5] = [1] * 5
    cl[11] = cl[900] = 3
    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)
    S = retrieval.AlignmentGallery(ti, tc, il, cl, n_images=Ni, img_start=0, img_step=5, precision=precision).scores()
    want = retrieval.rank_both_directions(S, Ni, k=50)
    got = retrieval.streaming_ranks(ti, tc, il, cl, Ni, img_start=0, img_step=5, precision=precision, block_images=block, k=50)
    for a, b, name in zip(got, want, ("ranks_i2t", "top1", "ranks_t2i", "top50")):
        np.testing.assert_array_equal(a, b, err_msg=name)
    if block == 1000:           # the same native call asked to write the matrix as well: one pass, identical scores and ranking
        got = retrieval.streaming_ranks(ti, tc, il, cl, Ni, img_start=0, img_step=5, precision=precision, block_images=block, k=50,
                                        keep_scores=True)
        for a, b, name in zip(got, want, ("ranks_i2t", "top1", "ranks_t2i", "top50")):
            np.testing.assert_array_equal(a, b, err_msg=name)
        assert torch.equal(got[4], S)


def test_oversized_gallery_takes_the_streaming_path(monkeypatch):
    """A score matrix above evaluation.STREAM_SCORE_BYTES is never materialised: i2t / t2i rank block by block and return
    what the dense path returns (golden retrieval set, fp32 mode: ranks, top-1, top-50, metrics)."""
    from aladin_b200 import evaluation as E, loss as L
    g = load_golden("retrieval")
    images = torch.from_numpy(np.repeat(g["images"], 5, axis=0))
    captions = torch.from_numpy(g["captions"])
    il, cl = g["img_lens"].tolist(), g["cap_lens"].tolist()
    crit = L.AlignmentContrastiveLoss(aggregation="MrSw")
    crit.precision = "fp32"
    E.clear_cache()
    dense = (E.i2t(images, captions, il, cl, return_ranks=True, sim_function=crit),
             E.t2i(images, captions, il, cl, return_ranks=True, sim_function=crit))
    assert E._cache["res"]["S"] is not None
    monkeypatch.setattr(E, "STREAM_SCORE_BYTES", 0)
    monkeypatch.setattr(E, "_stream_block", lambda Nc: 64)
    E.clear_cache()
    stream = (E.i2t(images, captions, il, cl, return_ranks=True, sim_function=crit),
              E.t2i(images, captions, il, cl, return_ranks=True, sim_function=crit))
    assert E._cache["res"]["S"] is None
    for (m0, (r0, l0)), (m1, (r1, l1)) in zip(dense, stream):
        assert m0 == m1
        np.testing.assert_array_equal(r0, r1)
        np.testing.assert_array_equal(l0, l1)
    np.testing.assert_array_equal(stream[0][1][0], g["ranks_i2t"])
