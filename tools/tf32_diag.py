import sys, numpy as np, torch
sys.path.insert(0,'.')
from aladin_b200 import retrieval, synth
from oracle import alad_oracle as O
for shape in [(37, 53, 35, 53, 128), (64, 320, 71, 71, 1024), (130, 90, 35, 53, 768)]:
    Bi, Bc, S_im, S_s, d = shape
    im, s, il, cl = synth.raw_batch(21, Bi, Bc, S_im, S_s, d, related=0.5)
    il[3] = 1
    ref = O.mrsw_scores(im, s, il, cl, acc64=True).astype(np.float64)
    out = {}
    for prec in ("tf32", "fp32", "bf16"):
        gal = retrieval.AlignmentGallery(torch.from_numpy(im).cuda(), torch.from_numpy(s).cuda(), il, cl, n_images=Bi, precision=prec)
        got = gal.scores().cpu().numpy().astype(np.float64)
        err = np.abs(got - ref)
        rel = err / np.maximum(np.abs(ref), 0.01 * np.abs(ref).max())
        out[prec] = (float(err.max()), float(rel.max()), float((got - ref).mean()))
    print(shape, "max|ref|", float(np.abs(ref).max()), {k: tuple(f"{x:.2e}" for x in v) for k, v in out.items()}, flush=True)
