"""aladin_b200 -- B200 (sm_100a) drop-ins for ALADIN's all-pairs cross-modal scoring path.

Mirrors the reference call sites (mesnico/ALADIN): ``alad.loss`` criteria, ``alad.evaluation``
i2t/t2i and ``alad.recall_auxiliary.compute_recall``.  All compute runs in hand-written CUDA
reached through the C ABI in include/alad_b200.h; there is no CPU or PyTorch fallback."""
from .scoring import alignment_scores, dot_scores, get_precision, set_precision  # noqa: F401

__all__ = ["alignment_scores", "dot_scores", "set_precision", "get_precision"]
