"""GPU parity tests for the fused training-loss call site (aladin_b200.alad_model: one native call per
direction for the three criteria of ALADModel.forward_loss, alad/alad_model.py:371-428) against the reference's
golden training step and against the per-criterion drop-ins."""
import numpy as np
import pytest
import torch

from conftest import CANCELLING_FLOOR, assert_scores_close, load_golden

pytestmark = pytest.mark.gpu


def cu(x, grad=False):
    return torch.tensor(np.asarray(x, np.float32), device="cuda", requires_grad=grad)


class _Logger:
    def __init__(self):
        self.rows = []

    def update(self, k, v, n=0):
        self.rows.append((k, v, n))


class _Model:
    """The attributes ALADModel.forward_loss touches (alad_model.py:263-292,371-428), with the drop-in criteria."""

    def __init__(self, loss_types, aggregation="MrSw", distill_mode="listnet", precision="fp32"):
        from aladin_b200 import loss as L
        self.config = {"training": {"loss-type": "-".join(loss_types)}}
        self.losses_types = list(loss_types)
        self.matching_criterion = L.ContrastiveLoss(margin=0.2, measure="dot", max_violation=True)
        self.alignment_criterion = L.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=True,
                                                              aggregation=aggregation)
        self.alignment_criterion.precision = precision
        self.distillation_loss = L.DistillationLoss(mode=distill_mode)
        self.logger = _Logger()
        self.original_calls = 0

    def forward_loss(self, img_emb, cap_emb, img_emb_set, cap_emb_seq, img_lengths, cap_lengths, reg_loss):
        """Per-criterion composition (what the reference method does with the drop-in criteria)."""
        self.original_calls += 1
        losses = {}
        ml, mm = self.matching_criterion(img_emb, cap_emb, return_similarity_mat=True)
        if "matching" in self.config["training"]["loss-type"]:
            losses["matching"] = ml
        al, ts = self.alignment_criterion(img_emb_set.permute(1, 0, 2), cap_emb_seq.permute(1, 0, 2), img_lengths, cap_lengths,
                                          return_similarity_mat=True)
        if "alignment" in self.losses_types:
            losses["alignment"] = al
        if "distillation" in self.losses_types:
            losses["distillation"] = self.distillation_loss(ts, mm)
        return losses


def test_fused_losses_match_reference_training_step():
    """Golden vectors produced by the unmodified ALADModel.forward_loss + weighting (tests/golden/make_golden.py)."""
    import aladin_b200
    from aladin_b200 import alad_model as AM
    g = load_golden("train_step")
    img_cls, cap_cls = cu(g["img_cls"], True), cu(g["cap_cls"], True)
    img_set, cap_seq = cu(g["img_set"], True), cu(g["cap_seq"], True)          # [S, B, d] like the model hands them over
    lm, la, ld, M, S = AM.train_losses(img_cls, cap_cls, img_set.permute(1, 0, 2), cap_seq.permute(1, 0, 2),
                                       g["img_len"].tolist(), g["cap_len"].tolist(), margin=0.2, max_violation=True,
                                       precision="fp32", precision_matching="fp32")
    loss = la * 1.0 + ld * 1.0 + lm * 0.1                                       # alad_model.py:449-453
    loss.backward()
    assert aladin_b200.get_precision() == "bf16"                                # the global mode is untouched
    assert_scores_close(S.cpu().numpy(), g["teacher_scores"], 1e-4, "teacher")
    assert_scores_close(M.cpu().numpy(), g["matching_mat"], 1e-4, "matching", floor=CANCELLING_FLOOR)
    np.testing.assert_allclose([lm.item(), la.item(), ld.item(), loss.item()],
                               [g["matching_loss"], g["alignment_loss"], g["distillation_loss"], g["loss"]], rtol=1e-4)
    np.testing.assert_allclose(img_cls.grad.cpu().numpy(), g["d_img_cls"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(cap_cls.grad.cpu().numpy(), g["d_cap_cls"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(img_set.grad.cpu().numpy(), g["d_img_set"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(cap_seq.grad.cpu().numpy(), g["d_cap_seq"], rtol=1e-3, atol=1e-5)
    assert img_set.grad.stride() == img_set.stride()                            # gradient in the [S, B, d] layout of the input
    assert not M.requires_grad and not S.requires_grad


def _batch(B, d, seed=3):
    from aladin_b200 import synth
    im, s, il, cl = synth.raw_batch(seed, B, B, 35, 53, d, related=0.6)
    r = np.random.RandomState(seed)
    icls = r.standard_normal((B, d)).astype(np.float32)
    ccls = (0.7 * icls + r.standard_normal((B, d))).astype(np.float32)
    icls /= np.linalg.norm(icls, axis=1, keepdims=True)
    ccls /= np.linalg.norm(ccls, axis=1, keepdims=True)
    return icls, ccls, im.transpose(1, 0, 2).copy(), s.transpose(1, 0, 2).copy(), il, cl


@pytest.mark.parametrize("types,pop_distill", [(("alignment", "matching", "distillation"), False),
                                               (("alignment", "matching", "distillation"), True),
                                               (("alignment", "distillation"), False),
                                               (("matching", "distillation"), False)])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_installed_forward_loss_equals_per_criterion_path(types, pop_distill, precision):
    """install() routes an eligible model through the fused call; losses are bit-identical to the per-criterion
    drop-ins (same kernels, same inputs), gradients agree to rounding (the fused backward sums the two
    contributions to dL/dM before its GEMMs).  pop_distill: the distillation loss is dropped before the
    distillation epoch (alad_model.py:441-443) -> no gradient reaches it."""
    from aladin_b200 import alad_model as AM
    icls, ccls, im, s, il, cl = _batch(96, 256)
    weights = {"alignment": 1.0, "distillation": 1.0, "matching": 0.1}

    def run(fused):
        class M(_Model):
            pass
        if fused:
            AM.install(M)
        m = M(types, precision=precision)
        leaves = [cu(icls, True), cu(ccls, True), cu(im, True), cu(s, True)]
        d = m.forward_loss(leaves[0], leaves[1], leaves[2], leaves[3], il, cl, None)
        if pop_distill:
            d.pop("distillation", None)
        sum(d[k] * weights[k] for k in d).backward()
        grads = [None if t.grad is None else t.grad.cpu().numpy() for t in leaves]
        return m, {k: v.item() for k, v in d.items()}, grads

    m_f, loss_f, g_f = run(True)
    m_r, loss_r, g_r = run(False)
    assert m_f.original_calls == 0 and m_r.original_calls == 1
    assert list(loss_f) == list(loss_r)                                          # same keys, same order
    assert loss_f == loss_r
    assert [k for k, _, _ in m_f.logger.rows] == [f"{k}_loss" for k in ("matching", "alignment", "distillation") if k in types]
    for k, v, n in m_f.logger.rows:
        assert n == 96 and (k.replace("_loss", "") not in loss_f or v == loss_f[k.replace("_loss", "")])
    for a, b, name in zip(g_f, g_r, ("img_emb", "cap_emb", "img_set", "cap_seq")):
        assert (a is None) == (b is None), name
        if a is not None:
            np.testing.assert_allclose(a, b, rtol=2e-4, atol=1e-6, err_msg=name)


def test_ineligible_configurations_use_the_original_method():
    from aladin_b200 import alad_model as AM
    icls, ccls, im, s, il, cl = _batch(16, 64)

    class M(_Model):
        pass
    AM.install(M)
    AM.install(M)                                                                # idempotent
    for kw in ({"loss_types": ("alignment", "matching"), "aggregation": "MrAVGw"},
               {"loss_types": ("alignment", "distillation"), "distill_mode": "mse"},
               {"loss_types": ("matching",)},
               {"loss_types": ("alignment", "matching", "regularizehidden")}):
        m = M(**kw)
        assert not AM.fused_eligible(m)
        if "regularizehidden" in kw["loss_types"]:
            continue
        out = m.forward_loss(cu(icls), cu(ccls), cu(im), cu(s), il, cl, None)
        assert m.original_calls == 1 and all(torch.isfinite(v).all() for v in out.values())


def test_forward_only_and_partial_requires_grad():
    from aladin_b200 import alad_model as AM
    icls, ccls, im, s, il, cl = _batch(40, 128)
    with torch.no_grad():
        a = AM.train_losses(cu(icls), cu(ccls), cu(im).permute(1, 0, 2), cu(s).permute(1, 0, 2), il, cl)
    x_i, x_s = cu(icls, True), cu(s, True)                                       # only two of the four inputs need gradients
    b = AM.train_losses(x_i, cu(ccls), cu(im).permute(1, 0, 2), x_s.permute(1, 0, 2), il, cl)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    (b[0] + b[1] + b[2]).backward()
    assert x_i.grad is not None and x_s.grad is not None and torch.isfinite(x_i.grad).all() and torch.isfinite(x_s.grad).all()
    # an empty batch is a no-op with zero losses
    e = AM.train_losses(cu(np.zeros((0, 128))), cu(np.zeros((0, 128))), cu(np.zeros((0, 35, 128))), cu(np.zeros((0, 53, 128))), [], [])
    assert [t.item() for t in e[:3]] == [0.0, 0.0, 0.0] and e[3].shape == (0, 0)
