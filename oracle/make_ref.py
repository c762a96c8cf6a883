"""TEST / BENCH INFRASTRUCTURE ONLY (same rules as alad_oracle.py: nothing under aladin_b200/ may import this).

Recipe for ``oracle/_ref``: the UNMODIFIED reference files of the hot path, copied byte for byte from the
reference checkout (default /root/reference) into the git-ignored directory ``oracle/_ref/alad`` so that they travel
to the GPU box with the snapshot (the checkout itself does not exist there).  Nothing is edited: the two
adaptations a CPU-only run needs (``Tensor.cuda`` as the identity, ``CUDA_VISIBLE_DEVICES=""`` for the stand-alone
arm) are applied at load time by ``oracle/ref_runner.py``.  A MANIFEST.json records the sha256 of every file.

    python oracle/make_ref.py [/path/to/reference]

``__graft_entry__.build()`` runs this when the checkout is present; ``bench.py`` then times these files as the CPU
arm (``cpu_baseline.kind = "reference"``) and falls back to the torch port (``"port"``) when they are absent."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["alad/__init__.py", "alad/loss.py", "alad/evaluation.py", "alad/recall_auxiliary.py", "alad/utils.py",
         "alad/evaluate_utils/dcg.py"]


def make(reference="/root/reference"):
    if not os.path.isdir(os.path.join(reference, "alad")):
        return None
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):
                continue
            raise FileNotFoundError(src)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference, "sha256": manifest}, f, indent=1)
    return DEST


if __name__ == "__main__":
    print(make(*sys.argv[1:2]))
