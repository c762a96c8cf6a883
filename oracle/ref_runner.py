"""TEST / BENCH INFRASTRUCTURE ONLY (same rules as alad_oracle.py: nothing under aladin_b200/ may import this).

Loads the UNMODIFIED reference files from ``oracle/_ref`` (see make_ref.py) and runs the reference's own public API
for the hot path on the host cores: ``alad.evaluation.i2t`` / ``t2i`` with the ``sim_function`` closure of
alad/test.py:259-263 over ``alad.loss.AlignmentContrastiveLoss(aggregation='MrSw')``.

The reference calls ``.cuda()`` unconditionally on that path (alad/evaluation.py:179,202,267,291); on a CPU run
``Tensor.cuda`` is made the identity for the duration of the call (restored afterwards, so the GPU arm of the same
process is unaffected) -- the reference's arithmetic is untouched."""
import contextlib
import importlib
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_mods = None


def available():
    return os.path.exists(os.path.join(REF, "alad", "evaluation.py")) and os.path.exists(os.path.join(REF, "alad", "loss.py"))


def load():
    """(alad.evaluation, alad.loss, alad.recall_auxiliary) of the reference copy."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
        import warnings
        sys.path.insert(0, REF)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")            # SyntaxWarning: "is" with a literal (recall_auxiliary.py)
                for name in [m for m in sys.modules if m == "alad" or m.startswith("alad.")]:
                    del sys.modules[name]
                ev = importlib.import_module("alad.evaluation")
                lo = importlib.import_module("alad.loss")
                ra = importlib.import_module("alad.recall_auxiliary")
        finally:
            sys.path.remove(REF)
        _mods = (ev, lo, ra)
    return _mods


@contextlib.contextmanager
def cpu_only():
    """Tensor.cuda / Module.cuda as the identity while the reference runs on the host cores."""
    import torch
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def sim_function():
    """The closure alad/test.py:259-263 hands to i2t / t2i."""
    import torch
    _, lo, _ = load()
    crit = lo.AlignmentContrastiveLoss(aggregation="MrSw")

    def alignment_sim_fn(img, cap, img_len, cap_len):
        with torch.no_grad():
            return crit(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

    return alignment_sim_fn


def run_sample(images, captions, img_lens, cap_lens, q, batches=5):
    """q query images through the reference's i2t (each against ALL captions) and q caption groups through its t2i
    (each against ALL images); returns (seconds, pairs scored, (i2t result, t2i result))."""
    ev, _, _ = load()
    fn = sim_function()
    with cpu_only(), contextlib.redirect_stderr(io.StringIO()):          # tqdm bars
        t0 = time.perf_counter()
        a = ev.i2t(images, captions, img_lens, cap_lens, npts=q, return_ranks=True, sim_function=fn, cap_batches=batches)
        b = ev.t2i(images, captions, img_lens, cap_lens, npts=q, return_ranks=True, sim_function=fn, im_batches=batches)
        dt = time.perf_counter() - t0
    Ni, Nc = images.shape[0] // 5, captions.shape[0]
    return dt, q * Nc + 5 * q * Ni, (a, b)
