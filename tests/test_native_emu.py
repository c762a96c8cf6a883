"""CPU-only: the `-m gpu` parity tests of the loss / scoring / ranking / evaluation drop-ins, UNCHANGED, on the "virtual B200" of
tests/conftest.py -- the Python host layer, the native compositions (alad_scores_fused, alad_train_losses_fwd/_bwd:
host bookkeeping, tile tables, workspace layout, launch order) and every CUDA-core kernel run from their real
sources on CPU tensors; only the tcgen05 scoring kernel is replaced by a CPU double of its documented contract
(tests/cuda_emu/mrsw_fwd_double.cpp).  The selection keeps to small shapes (one host thread per CUDA thread)."""
import pytest

import test_gpu_distill_modes as TD
import test_gpu_losses as TL
import test_gpu_ranking as TR
import test_gpu_retrieval as TE
import test_gpu_scoring as TS
import test_gpu_train_step as TT
import test_gpu_zz_scan_sentences as TZ
import test_gpu_zzz_ranking_consumers as TC

CASES = [
    (TS.test_pack_tokens_matches_normalize, {}),
    (TS.test_gemm_epilogue_is_bit_exact_on_exact_inputs, dict(Ni=7, Nc=33, d=64)),
    (TS.test_alignment_scores_golden, dict(precision="bf16")),
    (TS.test_alignment_scores_golden, dict(precision="fp32")),
    (TS.test_all_pooling_modes_golden, dict(precision="fp32")),
    (TS.test_mrsw_packed_is_bit_exact_on_exact_inputs, {}),
    (TS.test_pooling_modes_vs_oracle_ragged, dict(agg="symm")),
    (TL.test_other_pooling_modes_gradients_vs_torch_autograd, dict(agg="MwSr")),
    (TT.test_ineligible_configurations_use_the_original_method, {}),
    (TL.test_triplet_golden, dict(key="mv", mv=True)),
    (TL.test_listnet_golden, {}),
    (TL.test_matching_golden, {}),
    (TL.test_alignment_loss_golden, dict(key="mv", mv=True)),
    (TL.test_alignment_loss_golden, dict(key="sum", mv=False)),
    (TL.test_alignment_dense_upstream_gradient_golden, {}),
    (TL.test_train_step_call_site_golden, {}),
    (TD.test_mse_golden, {}),
    (TD.test_contrastive_golden, dict(margin=0.2)),
    (TD.test_ordinal_golden, dict(margin=0.2, thr=0.1, stride=3)),
    (TD.test_order_sim_golden_and_gradient, {}),
    (TD.test_cosine_measure_gradient_golden, dict(key="cosine_mv", mv=True)),
    (TD.test_pooling_mode_gradients_golden, dict(agg="symm")),
    (TD.test_pooling_mode_gradients_golden, dict(agg="mean")),
    (TR.test_rank_kernels_reproduce_reference_golden, {}),
    (TR.test_sharded_ranking_equals_single_shard, {}),
    (TR.test_col_topk_select_strided_view, {}),
    (TR.test_rank_fused_reproduces_reference_golden, {}),
    (TR.test_rank_fused_equals_one_purpose_kernels, dict(Ni=300, Nc=200, q_rows=260, q_cols=120, k=10, img_off=2, kind="given")),
    (TE.test_i2t_t2i_alignment_golden, dict(precision="fp32")),
    (TE.test_arbitrary_callable_sim_function, {}),
    (TC.test_recall_1k_5fold_small_folds_and_missing_folds, {}),
    (TC.test_ndcg_scorer_hooks_receive_the_reference_order, {}),
    (TT.test_fused_losses_match_reference_training_step, {}),
    (TZ.test_scores_golden, dict(precision="fp32")),
    (TZ.test_degenerate_lengths_like_reference, {}),
    (TZ.test_gradients_golden_full_length_images, dict(precision="fp32", tol=1e-3)),
    (TZ.test_gradients_ragged_vs_oracle_and_reference_where_finite, {}),
]


@pytest.mark.parametrize("fn,kwargs", CASES, ids=[f"{f.__module__}.{f.__name__}[{','.join(map(str, k.values()))}]" for f, k in CASES])
def test_gpu_test_body_on_the_virtual_device(virtual_b200, monkeypatch, capsys, fn, kwargs):
    import inspect
    if "monkeypatch" in inspect.signature(fn).parameters:
        kwargs = dict(kwargs, monkeypatch=monkeypatch)
    if "capsys" in inspect.signature(fn).parameters:
        kwargs = dict(kwargs, capsys=capsys)
    fn(**kwargs)


@pytest.mark.parametrize("agg", ["symm", "scan-sentences"])
def test_block_mode_of_i2t_on_a_small_gallery(virtual_b200, agg):
    """evaluation.i2t with a closure over the drop-in criterion in a pooling mode other than 'MrSw': one criterion
    call on the whole gallery (the GPU test of the same path also runs the per-query loop, which is too slow for
    the emulator); scores and ranks against the oracle."""
    import numpy as np
    import torch
    from aladin_b200 import evaluation, loss as L, synth
    from oracle import alad_oracle as O
    images, captions, il, cl = synth.eval_containers(3, 8, 14, 32, max_regions=9, max_words=10)
    crit = L.AlignmentContrastiveLoss(aggregation=agg)
    crit.precision = "fp32"
    calls = []
    orig = crit.forward
    crit.forward = lambda *a, **k: (calls.append(a[0].shape[0]), orig(*a, **k))[1]
    evaluation.clear_cache()
    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)      # kept alive: the cached block dies with its inputs
    m, (ranks, top1) = evaluation.i2t(ti, tc, il, cl, return_ranks=True,
                                      sim_function=lambda im, cap, a, b: crit(im, cap, a, b, return_loss=False,
                                                                              return_similarity_mat=True), cap_batches=2)
    assert calls == [1, 1, 8]                              # probe of the closure (closure, criterion), then one block call
    ref = (O.alignment_scores_small(images[0::5], captions, il[0::5], cl, agg) if agg != "scan-sentences"
           else O.scan_scores(images[0::5], captions, il[0::5], cl))
    S = evaluation._cache["res"]["S"].numpy()
    np.testing.assert_allclose(S, ref, rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(ranks, O.i2t_ranks(ref)[0])
    assert evaluation._cache["key"][-2] == f"block:{agg}"


@pytest.mark.parametrize("block,keep", [(16, False), (100, True)])
def test_native_streaming_retrieval_on_a_small_gallery(virtual_b200, block, keep):
    """alad_mrsw_retrieval (the native block-by-block composition behind retrieval.streaming_ranks): ranks, top-1 and top-k
    equal the ranking of the dense matrix -- several blocks with a ragged last one, and one block in the variant that also
    writes the matrix.  (The GPU test of the same function runs a 333-image gallery; too slow for the emulator.)"""
    import numpy as np
    import torch
    from aladin_b200 import retrieval, synth
    Ni, k = 41, 10
    images, captions, il, cl = synth.eval_containers(5, Ni, 24, 64, max_regions=12, max_words=14, alpha=0.3)
    il[5 * 3:5 * 3 + 5] = [1] * 5                  # an image without regions
    cl[7] = cl[101] = 3                            # captions without words
    ti, tc = torch.from_numpy(images), torch.from_numpy(captions)
    S = retrieval.AlignmentGallery(ti, tc, il, cl, n_images=Ni, img_start=0, img_step=5, precision="bf16").scores()
    want = retrieval.rank_both_directions(S, Ni, k=k)
    got = retrieval.streaming_ranks(ti, tc, il, cl, Ni, img_start=0, img_step=5, precision="bf16", block_images=block, k=k,
                                    keep_scores=keep)
    for a, b, name in zip(got, want, ("ranks_i2t", "top1", "ranks_t2i", "topk")):
        np.testing.assert_array_equal(a, b, err_msg=name)
    if keep:
        assert torch.equal(got[4], S)
