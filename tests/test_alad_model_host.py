"""CPU-only: host logic of the fused forward_loss call site (aladin_b200.alad_model): eligibility rules and the
install() routing.  No compute calls (the fused path itself is covered by tests/test_gpu_train_step.py)."""
from aladin_b200 import alad_model as AM, loss as L


class _Model:
    def __init__(self, loss_types, aggregation="MrSw", distill_mode="listnet", measure="dot"):
        self.config = {"training": {"loss-type": "-".join(loss_types)}}
        self.losses_types = list(loss_types)
        self.matching_criterion = L.ContrastiveLoss(margin=0.2, measure=measure, max_violation=True)
        self.alignment_criterion = L.AlignmentContrastiveLoss(margin=0.2, measure=measure, max_violation=True,
                                                              aggregation=aggregation)
        if "distillation" in loss_types:
            self.distillation_loss = L.DistillationLoss(mode=distill_mode)

    def forward_loss(self, *args):
        return "original"


def test_eligibility_follows_the_shipped_configurations():
    assert AM.fused_eligible(_Model(("alignment", "matching", "distillation")))      # configs/alad-alignment-and-matching-distill*.yaml
    assert AM.fused_eligible(_Model(("alignment", "matching")))
    assert AM.fused_eligible(_Model(("alignment",)))
    assert AM.fused_eligible(_Model(("matching", "distillation")))
    assert not AM.fused_eligible(_Model(("matching",)))                               # nothing to fuse: one criterion
    assert not AM.fused_eligible(_Model(("alignment", "matching"), aggregation="symm"))
    assert not AM.fused_eligible(_Model(("alignment", "matching"), measure="cosine"))
    assert not AM.fused_eligible(_Model(("alignment", "distillation"), distill_mode="ordinal"))
    assert not AM.fused_eligible(_Model(("alignment", "matching", "selfaggregation")))
    assert not AM.fused_eligible(_Model(("alignment", "matching", "entropy")))
    assert not AM.fused_eligible(_Model(("alignment", "matching", "regularizehidden")))

    class Foreign(_Model):
        pass
    m = Foreign(("alignment", "matching"))
    m.alignment_criterion = object()                                                  # not an aladin_b200 criterion
    assert not AM.fused_eligible(m)


def test_install_routes_ineligible_models_to_the_original_and_is_idempotent():
    class M(_Model):
        pass
    AM.install(M)
    first = M.forward_loss
    AM.install(M)
    assert M.forward_loss is first and first._alad_b200_fused
    assert M(("matching",)).forward_loss(None, None, None, None, [], [], None) == "original"
    assert M(("alignment", "matching"), aggregation="MwSr").forward_loss(None, None, None, None, [], [], None) == "original"
