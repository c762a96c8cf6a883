"""CPU-only host logic of aladin_b200.evaluation: the score block computed by i2t is reused by t2i on the SAME input
objects (alad/test.py:271,275 call both on the same tensors) and never for other objects, even if address, shape,
version counter and lengths all recur (next epoch's validation embeddings).  The CUDA entry points are replaced by
torch / oracle doubles: this tests the caching rule, not the kernels."""
import numpy as np
import torch

from oracle import alad_oracle as O


def _install_doubles(monkeypatch, calls):
    from aladin_b200 import evaluation

    def dot_scores(im, s, precision=None, **kw):
        calls.append("scores")
        return im.float() @ s.float().t()

    def rank_both_directions(S, npts, img_off=0, n_images_total=None, k=50, group=None, **kw):
        S = S.numpy()
        ri, t1 = O.i2t_ranks(S)
        rt, tk = O.t2i_ranks(S, k)
        return ri, t1, rt, tk

    monkeypatch.setattr(evaluation.scoring, "dot_scores", dot_scores)
    monkeypatch.setattr(evaluation.retrieval, "rank_both_directions", rank_both_directions)
    evaluation.clear_cache()
    return evaluation


def _containers(seed, Ni=60, S=5, d=16):
    r = np.random.RandomState(seed)
    base = r.standard_normal((Ni, S, d)).astype(np.float32)
    images = torch.from_numpy(np.repeat(base, 5, axis=0))
    captions = torch.from_numpy((np.repeat(base, 5, axis=0) + 0.5 * r.standard_normal((5 * Ni, S, d))).astype(np.float32))
    return images, captions, [S] * (5 * Ni), [S] * (5 * Ni)


def test_score_block_is_shared_by_i2t_and_t2i_on_the_same_objects(monkeypatch):
    calls = []
    ev = _install_doubles(monkeypatch, calls)
    images, captions, il, cl = _containers(0)
    m1, (ranks, top1) = ev.i2t(images, captions, il, cl, return_ranks=True)
    m2, (ranks_t, top50) = ev.t2i(images, captions, il, cl, return_ranks=True)
    assert calls == ["scores"]
    S = images[0::5][:, 0, :].numpy() @ captions[:, 0, :].numpy().T
    np.testing.assert_array_equal(ranks, O.i2t_ranks(S)[0])
    np.testing.assert_array_equal(ranks_t, O.t2i_ranks(S)[0])
    assert len(m1) == 7 and len(m2) == 7 and top50.shape == (300, 50)


def test_cache_misses_for_other_objects_with_the_same_key(monkeypatch):
    calls = []
    ev = _install_doubles(monkeypatch, calls)
    images, captions, il, cl = _containers(1)
    r_old = ev.i2t(images, captions, il, cl)
    key_old = ev._cache["key"]
    # a different epoch: new tensors; force the recorded key to equal the new inputs' key (recurring address / version)
    images2, captions2, _, _ = _containers(2)
    ev._cache["key"] = ev._key(images2, captions2, ev.retrieval.lens_key(il, cl), "global", ev.scoring.get_precision())
    r_new = ev.i2t(images2, captions2, il, cl)
    assert calls == ["scores", "scores"] and key_old != ev._cache["key"]
    assert r_new != r_old
    # in-place modification of the same objects bumps the version counter -> recomputed as well
    images2.mul_(-1.0)
    ev.i2t(images2, captions2, il, cl)
    assert calls == ["scores"] * 3
    # dead inputs never hit, and the cached block (0.5 GB of HBM at COCO-5k) is released with them
    assert ev._cache["refs"][0]() is images2
    del images2, captions2
    assert not ev._cache
