"""CPU (host-thread CUDA emulator): the two-stage retrieval path -- stage-1 shortlists, the pair-tile table built on
the "device" (alad_pairtile_build: union bitmaps, scan, slot records), the pair-list scoring contract (CPU double of
the tcgen05 kernel), list gather + re-rank -- against the oracle composition of the reference's pieces."""
import numpy as np
import pytest
import torch

import test_gpu_two_stage as T2


def _tiles(details, slots):
    n = int(details["n_ptiles"].item())
    return n


@pytest.mark.parametrize("K,l2_bytes", [(7, 48 << 20), (100, 48 << 20), (12, 1 << 13)])
def test_two_stage_pair_list_path_on_the_emulator(virtual_b200, K, l2_bytes, monkeypatch):
    from aladin_b200 import synth, two_stage
    monkeypatch.setattr(two_stage, "L2_BLOCK_BYTES", l2_bytes)      # 8 KB: the 36 images go in two blocks of 32
    images, captions, il, cl = synth.eval_containers(43, 36, 14, 32, max_regions=9, max_words=12, alpha=0.3)
    il[5:10] = [1] * 5                      # an image without scored regions: its scores are 0
    cl[7] = 3                               # a caption without scored words: its column is 0
    out, det = two_stage.two_stage_retrieval(torch.from_numpy(images), torch.from_numpy(captions), il, cl, shortlist=K,
                                             precision="fp32", return_details=True)
    ei, et = T2._oracle_two_stage(images, captions, il, cl, K)
    np.testing.assert_array_equal(det["ranks_i2t"], ei)
    np.testing.assert_array_equal(det["ranks_t2i"], et)
    # every shortlisted pair was scored by exactly the tiles of the table: compare the gathered scores with the oracle
    from oracle import alad_oracle as O
    S_ref = O.mrsw_scores(images[0::5], captions, il[0::5], cl, acc64=True)
    ids = det["short_t2i"].numpy()
    got = det["scores_t2i"].numpy()
    ref = np.take_along_axis(S_ref.T, np.maximum(ids, 0), axis=1)
    np.testing.assert_allclose(got[ids >= 0], ref[ids >= 0], rtol=1e-4, atol=1e-5)
    # the re-ranked order is the shortlist sorted by alignment score (ties: list position descending)
    order = det["order_t2i"].numpy()
    for c in (0, 11, 90):
        valid = ids[c][ids[c] >= 0]
        exp = sorted(valid.tolist(), key=lambda i: (-got[c][list(ids[c]).index(i)], -list(ids[c]).index(i)))
        assert order[c][:len(valid)].tolist() == exp
    assert int(det["n_ptiles"].item()) > 0


def test_caption_groups_host_helper():
    from aladin_b200 import two_stage
    nw = np.array([50, 50, 50, 0, 100, 28, 1, 128, 0, 5], np.int32)
    n_g, row0, cap_lo, cap_group, word_box = two_stage._caption_groups(nw)
    assert word_box % 8 == 0 and 8 <= word_box <= 128
    assert n_g == 6
    assert cap_lo.tolist() == [0, 2, 4, 6, 7, 9, 10]
    assert row0.tolist() == [0, 100, 150, 278, 279, 407]
    assert cap_group.tolist() == [0, 0, 1, -1, 2, 2, 3, 4, -1, 5]
    with pytest.raises(Exception):
        two_stage._caption_groups(np.array([129], np.int32))
