"""The oracle is only trusted once it reproduces the reference's own outputs.

Golden vectors: tests/golden/*.npz, produced by tests/golden/make_golden.py from the
unmodified reference (alad/loss.py, alad/evaluation.py, alad/recall_auxiliary.py)."""
import numpy as np
import pytest

from oracle import alad_oracle as O
from conftest import load_golden


def test_alignment_scores_all_modes_match_reference():
    g = load_golden("alignment_scores")
    im_len, s_len = g["im_len"].tolist(), g["s_len"].tolist()
    for agg in O.AGGREGATIONS:
        S = O.alignment_scores_small(g["im"], g["s"], im_len, s_len, agg)
        ref = g["S_" + agg]
        # MrAVGw divides by nw == 0 for one caption -> nan/inf in the reference too
        np.testing.assert_allclose(S, ref, rtol=2e-5, atol=2e-6, equal_nan=True, err_msg=agg)


def test_mrsw_fast_and_scalar_match_reference():
    g = load_golden("alignment_scores")
    im_len, s_len = g["im_len"].tolist(), g["s_len"].tolist()
    ref = g["S_MrSw"]
    np.testing.assert_allclose(O.mrsw_scores(g["im"], g["s"], im_len, s_len, chunk=2), ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(O.mrsw_scores(g["im"], g["s"], im_len, s_len, acc64=True), ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(O.mrsw_scores_scalar(g["im"], g["s"], im_len, s_len), ref, rtol=2e-5, atol=2e-6)
    # reference edge cases: im_len==1 -> zero row, s_len==3 -> zero column (SURVEY A.5 iv)
    assert np.all(ref[2] == 0) and np.all(ref[:, 1] == 0)


def test_valid_count_python_slice_semantics():
    assert O.valid_count(5, 8) == 5
    assert O.valid_count(9, 8) == 8
    assert O.valid_count(0, 8) == 0
    assert O.valid_count(-1, 8) == 7        # mask[-1:] masks only the last slot
    assert O.valid_count(-9, 8) == 0


@pytest.mark.parametrize("key,mv", [("mv", True), ("sum", False)])
def test_alignment_loss_and_grads_match_reference(key, mv):
    g = load_golden("alignment_loss")
    im = np.transpose(g["im_sbd"], (1, 0, 2))
    s = np.transpose(g["s_sbd"], (1, 0, 2))
    im_len, s_len = g["im_len"].tolist(), g["s_len"].tolist()
    S = O.mrsw_scores(im, s, im_len, s_len)
    np.testing.assert_allclose(S, g[f"S_{key}"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(O.triplet_loss(S, 0.2, mv), g[f"loss_{key}"], rtol=1e-5)
    G = O.triplet_grad(g[f"S_{key}"], 0.2, mv)
    d_im, d_s = O.mrsw_backward(im, s, im_len, s_len, G)
    np.testing.assert_allclose(np.transpose(d_im, (1, 0, 2)), g[f"dim_{key}"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(np.transpose(d_s, (1, 0, 2)), g[f"ds_{key}"], rtol=1e-4, atol=2e-6)


def test_alignment_dense_backward_matches_reference():
    g = load_golden("alignment_loss")
    im = np.transpose(g["im_sbd"], (1, 0, 2))
    s = np.transpose(g["s_sbd"], (1, 0, 2))
    d_im, d_s = O.mrsw_backward(im, s, g["im_len"].tolist(), g["s_len"].tolist(), g["Gup"])
    np.testing.assert_allclose(np.transpose(d_im, (1, 0, 2)), g["dim_dense"], rtol=1e-4, atol=3e-6)
    np.testing.assert_allclose(np.transpose(d_s, (1, 0, 2)), g["ds_dense"], rtol=1e-4, atol=3e-6)
    # dropped slots (image slot 0, caption slot 0 and the last two) never receive gradient
    assert np.all(d_im[:, 0] == 0) and np.all(d_s[:, 0] == 0) and np.all(d_s[:, -2:] == 0)


def test_matching_scores_and_loss_match_reference():
    g = load_golden("matching")
    for measure in ("dot", "cosine"):
        im = g["im"] * (1.0 if measure == "dot" else 2.5)
        S = O.dot_scores(im, g["s"]) if measure == "dot" else O.cosine_scores(im, g["s"])
        for key, mv in (("mv", True), ("sum", False)):
            k = f"{measure}_{key}"
            np.testing.assert_allclose(S, g[f"S_{k}"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(O.triplet_loss(S, 0.2, mv), g[f"loss_{k}"], rtol=1e-5)
            if measure == "dot":
                G = O.triplet_grad(g[f"S_{k}"], 0.2, mv)
                np.testing.assert_allclose(G @ g["s"], g[f"dim_{k}"], rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(G.T @ im, g[f"ds_{k}"], rtol=1e-5, atol=1e-6)


def test_triplet_and_listnet_match_reference():
    g = load_golden("triplet_listnet")
    for key, mv in (("mv", True), ("sum", False)):
        np.testing.assert_allclose(O.triplet_loss(g["S"], 0.2, mv), g[f"loss_{key}"], rtol=1e-6)
        np.testing.assert_array_equal(O.triplet_grad(g["S"], 0.2, mv), g[f"G_{key}"])
    np.testing.assert_allclose(O.listnet_loss(g["T"], g["M"]), g["listnet_loss"], rtol=1e-5)
    np.testing.assert_allclose(O.listnet_grad(g["T"], g["M"]), g["listnet_dM"], rtol=2e-4, atol=1e-7)
    assert bool(g["listnet_dT_is_none"])   # teacher is detached (loss.py:370)


def _containers(g):
    return np.repeat(g["images"], 5, axis=0), g["captions"], g["img_lens"].tolist(), g["cap_lens"].tolist()


def test_retrieval_alignment_matches_reference():
    g = load_golden("retrieval")
    images, captions, il, cl = _containers(g)
    S = O.mrsw_scores(images[0::5], captions, il[0::5], cl)
    np.testing.assert_allclose(S, g["S_full"], rtol=2e-5, atol=2e-6)
    ri, top1 = O.i2t_ranks(S)
    rt, top50 = O.t2i_ranks(S)
    np.testing.assert_array_equal(ri, g["ranks_i2t"])
    np.testing.assert_array_equal(top1, g["top1"])
    np.testing.assert_array_equal(rt, g["ranks_t2i"])
    np.testing.assert_array_equal(top50, g["top50"])
    np.testing.assert_allclose(O.recall_metrics(ri), g["m_i2t"][:5])
    np.testing.assert_allclose(O.recall_metrics(rt), g["m_t2i"][:5])


def test_retrieval_loops_match_reference():
    g = load_golden("retrieval")
    images, captions, il, cl = _containers(g)
    m, (ranks, top1) = O.i2t(images, captions, il, cl, cap_batches=5)
    np.testing.assert_allclose(m, g["m_i2t"])
    np.testing.assert_array_equal(ranks, g["ranks_i2t"])
    m, (ranks, top50) = O.t2i(images, captions, il, cl, im_batches=5)
    np.testing.assert_allclose(m, g["m_t2i"])
    np.testing.assert_array_equal(top50, g["top50"])
    # slot-0 (global vector) path, sim_function=None
    m, (ranks, top1) = O.i2t(images, captions, il, cl, use_alignment=False)
    np.testing.assert_allclose(m, g["g_i2t"])
    np.testing.assert_array_equal(top1, g["gtop1"])
    m, (ranks, top50) = O.t2i(images, captions, il, cl, use_alignment=False)
    np.testing.assert_allclose(m, g["g_t2i"])
    np.testing.assert_array_equal(ranks, g["granks_t2i"])
    np.testing.assert_allclose(O.compute_recall(images[:, 0, :], captions[:, 0, :]), g["compute_recall"])


def test_train_step_composite_matches_reference():
    """alad_model.py:377-405 + :445-453 restated with the oracle pieces."""
    g = load_golden("train_step")
    im = np.transpose(g["img_set"], (1, 0, 2))
    s = np.transpose(g["cap_seq"], (1, 0, 2))
    il, cl = g["img_len"].tolist(), g["cap_len"].tolist()
    M = O.dot_scores(g["img_cls"], g["cap_cls"])
    T = O.mrsw_scores(im, s, il, cl)
    np.testing.assert_allclose(M, g["matching_mat"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(T, g["teacher_scores"], rtol=2e-5, atol=2e-6)
    lm, la, ld = O.triplet_loss(M, 0.2, True), O.triplet_loss(T, 0.2, True), O.listnet_loss(T, M)
    np.testing.assert_allclose([lm, la, ld], [g["matching_loss"], g["alignment_loss"], g["distillation_loss"]], rtol=2e-5)
    np.testing.assert_allclose(la + ld + 0.1 * lm, g["loss"], rtol=2e-5)
    GM = 0.1 * O.triplet_grad(M, 0.2, True) + O.listnet_grad(T, M)
    np.testing.assert_allclose(GM @ g["cap_cls"], g["d_img_cls"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(GM.T @ g["img_cls"], g["d_cap_cls"], rtol=2e-4, atol=2e-6)
    d_im, d_s = O.mrsw_backward(im, s, il, cl, O.triplet_grad(T, 0.2, True))
    np.testing.assert_allclose(np.transpose(d_im, (1, 0, 2)), g["d_img_set"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(np.transpose(d_s, (1, 0, 2)), g["d_cap_seq"], rtol=1e-4, atol=2e-6)


# ------------------------------------------------------------------ remaining distillation modes
def test_distill_mse_matches_reference():
    g = load_golden("distill_modes")
    loss, dM, dwb = O.distill_mse(g["T"], g["M"], g["mse_wb"])
    np.testing.assert_allclose(loss, g["mse_loss"], rtol=1e-6)
    np.testing.assert_allclose(dM, g["mse_dM"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(dwb, g["mse_dwb"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("margin", [0.2, 0.05])
def test_distill_contrastive_matches_reference(margin):
    g = load_golden("distill_modes")
    loss, dM = O.distill_contrastive(g["T"], g["M"], margin)
    k = f"contrastive_m{margin}"
    np.testing.assert_allclose(loss, g[k + "_loss"], rtol=1e-6)
    np.testing.assert_array_equal(dM, g[k + "_dM"])                  # integer-valued gradient
    # the reference zeroes the caller's teacher diagonal in place (alad/loss.py:400)
    assert np.all(np.diag(g[k + "_T_after"]) == 0)


@pytest.mark.parametrize("margin,thr,stride", [(0.2, 0.1, 3), (0.1, 0.5, 1)])
def test_distill_ordinal_matches_reference(margin, thr, stride):
    g = load_golden("distill_modes")
    loss, dM = O.distill_ordinal(g["T"], g["M"], margin, thr, stride)
    k = f"ordinal_m{margin}_t{thr}_s{stride}"
    np.testing.assert_allclose(loss, g[k + "_loss"], rtol=1e-6)
    np.testing.assert_allclose(dM, g[k + "_dM"], rtol=1e-6, atol=1e-9)


def test_distill_ordinal_empty_selection_is_nan_like_reference():
    g = load_golden("distill_modes")
    loss, _ = O.distill_ordinal(g["T"], g["M"], 0.2, 100.0, 3)
    assert np.isnan(loss) and np.isnan(g["ordinal_m0.2_t100.0_s3_loss"])


def test_order_sim_matches_reference():
    g = load_golden("distill_modes")
    np.testing.assert_allclose(O.order_scores(g["order_im"], g["order_s"]), g["order_S"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("agg", ["sum", "mean"])
def test_pooled_sum_backward_matches_reference(agg):
    g = load_golden("pooled_grads")
    d_im, d_s = O.pooled_sum_backward(g["im"], g["s"], g["im_len"].tolist(), g["s_len"].tolist(), g["Gup"], mean=agg == "mean")
    np.testing.assert_allclose(d_im, g["dim_" + agg], rtol=1e-4, atol=1e-6 * np.abs(g["dim_" + agg]).max())
    np.testing.assert_allclose(d_s, g["ds_" + agg], rtol=1e-4, atol=1e-6 * np.abs(g["ds_" + agg]).max())


@pytest.mark.parametrize("key", ["cosine_mv", "cosine_sum"])
def test_cosine_backward_matches_reference(key):
    g = load_golden("matching")
    im = g["im"] * 2.5
    G = O.triplet_grad(g["S_" + key], 0.2, key.endswith("mv"))
    d_im, d_s = O.cosine_backward(im, g["s"], G)
    np.testing.assert_allclose(d_im, g["dim_" + key], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(d_s, g["ds_" + key], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------ torch-CPU port used as the timed CPU baseline
def test_torch_port_matches_reference_golden():
    import torch
    from oracle import alad_torch_port as TP
    g = load_golden("alignment_scores")
    S = TP.alignment_scores(torch.from_numpy(g["im"]), torch.from_numpy(g["s"]), g["im_len"].tolist(), g["s_len"].tolist())
    np.testing.assert_allclose(S.numpy(), g["S_MrSw"], rtol=2e-5, atol=2e-6)
    r = load_golden("retrieval")
    images = np.repeat(r["images"], 5, axis=0)
    il, cl = r["img_lens"].tolist(), r["cap_lens"].tolist()
    m_i, (ranks_i, top1) = TP.i2t(images, r["captions"], il, cl, cap_batches=5)
    m_t, (ranks_t, top50) = TP.t2i(images, r["captions"], il, cl, im_batches=5)
    np.testing.assert_array_equal(ranks_i, r["ranks_i2t"])
    np.testing.assert_array_equal(top1, r["top1"])
    np.testing.assert_array_equal(ranks_t, r["ranks_t2i"])
    np.testing.assert_array_equal(top50, r["top50"])
    np.testing.assert_allclose(m_i[:5], r["m_i2t"][:5])
    np.testing.assert_allclose(m_t[:5], r["m_t2i"][:5])


def test_scan_sentences_scores_match_reference():
    """aggregation 'scan-sentences' (alad/loss.py:136-149): ragged, training layout, degenerate lengths."""
    g = load_golden("scan_sentences")
    S = O.scan_scores(g["a_im"], g["a_s"], g["a_im_len"].tolist(), g["a_s_len"].tolist())
    np.testing.assert_allclose(S, g["a_S"], rtol=2e-5, atol=2e-6)
    im, s = np.transpose(g["b_im"], (1, 0, 2)), np.transpose(g["b_s"], (1, 0, 2))
    S = O.scan_scores(im, s, g["b_im_len"].tolist(), g["b_s_len"].tolist())
    np.testing.assert_allclose(S, g["b_S"], rtol=2e-5, atol=2e-6)
    for key, mv in (("mv", True), ("sum", False)):
        np.testing.assert_allclose(O.triplet_loss(S, 0.2, mv), g[f"b_loss_{key}"], rtol=1e-5)
    # an image without valid regions scores 0, a caption without valid words NaN (reference: softmax of -inf)
    S = O.scan_scores(g["c_im"], g["c_s"], g["c_im_len"].tolist(), g["c_s_len"].tolist(), acc64=False)
    ref = g["c_S"]
    assert np.array_equal(np.isnan(S), np.isnan(ref)) and np.isnan(ref[0, 1]) and np.all(ref[1] == 0)
    np.testing.assert_allclose(S, ref, rtol=2e-5, atol=2e-6, equal_nan=True)


def test_scan_sentences_backward_matches_reference_where_it_is_finite():
    g = load_golden("scan_sentences")
    il, sl = g["d_im_len"].tolist(), g["d_s_len"].tolist()
    np.testing.assert_allclose(O.scan_scores(g["d_im"], g["d_s"], il, sl), g["d_S"], rtol=2e-5, atol=2e-6)
    d_im, d_s = O.scan_backward(g["d_im"], g["d_s"], il, sl, g["d_Gup"])
    np.testing.assert_allclose(d_im, g["d_dim"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(d_s, g["d_ds"], rtol=1e-4, atol=2e-6)
    G = O.triplet_grad(g["d_S"], 0.2, True)
    d_im, d_s = O.scan_backward(g["d_im"], g["d_s"], il, sl, G)
    np.testing.assert_allclose(d_im, g["d_dim_loss"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(d_s, g["d_ds_loss"], rtol=1e-4, atol=2e-6)
    # ragged images: the reference's gradient is NaN on every scored word slot and on the masked region slots;
    # where it is finite (images without masked slots) it equals the oracle's
    il, sl = g["a_im_len"].tolist(), g["a_s_len"].tolist()
    d_im, d_s = O.scan_backward(g["a_im"], g["a_s"], il, sl, g["a_Gup"])
    ref = g["a_dim"]
    assert np.isnan(g["a_ds"][:, 1:-2]).all() and np.isfinite(d_im).all() and np.isfinite(d_s).all()
    full = [i for i, l in enumerate(il) if l == g["a_im"].shape[1]]
    assert full and np.isfinite(ref[full]).all()
    np.testing.assert_allclose(d_im[full], ref[full], rtol=1e-4, atol=2e-6)


def test_scan_fragile_pairs_flags_the_discontinuity_of_the_reference_op():
    """relu -> F.normalize over the regions (alad/loss.py:137-138): the golden gallery has exactly three pairs with a
    word column whose only non-negative cosine is ~0; perturbing the cosines at the 1e-5 level (what separates two
    fp32-grade GEMMs) moves one of them by 2.5e-2 and every other pair by < 2e-5."""
    g = load_golden("retrieval")
    il, cl = g["img_lens"].tolist()[0::5], g["cap_lens"].tolist()
    fr = O.scan_fragile_pairs(g["images"], g["captions"], il, cl)
    assert sorted(map(tuple, np.argwhere(fr).tolist())) == [(4, 45), (8, 260), (38, 122)]
