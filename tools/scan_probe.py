"""Timing of aggregation 'scan-sentences' (alad/loss.py:136-149) through the drop-in criterion at the training
shapes of BASELINE configs 1 and 4 (34 regions x 50 words, d = 1024): forward and forward + backward with the
hardest-negative hinge, CUDA events, against MrSw on the same inputs.  Run on the GPU box:
    python tools/scan_probe.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
import aladin_b200  # noqa: E402,F401
from aladin_b200 import loss as L, synth  # noqa: E402


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = []
    # `python tools/scan_probe.py one B` = one forward + backward of scan-sentences at batch B (for ncu launch lists)
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        B = int(sys.argv[2])
        im, s, il, cl = synth.raw_batch(9, B, B, 35, 53, 1024, related=0.6)
        img = torch.tensor(im, device="cuda", requires_grad=True)
        cap = torch.tensor(s, device="cuda", requires_grad=True)
        crit = L.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=True, aggregation="scan-sentences")
        crit.precision = "bf16"
        crit(img, cap, il, cl).backward()
        torch.cuda.synchronize()
        return
    for B, iters in ((128, 20), (512, 5)):
        im, s, il, cl = synth.raw_batch(9, B, B, 35, 53, 1024, related=0.6)
        img = torch.tensor(im, device="cuda", requires_grad=True)
        cap = torch.tensor(s, device="cuda", requires_grad=True)
        for agg in ("MrSw", "scan-sentences"):
            for precision in ("bf16", "fp32"):
                crit = L.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=True, aggregation=agg)
                crit.precision = precision

                def fwd():
                    with torch.no_grad():
                        return crit(img, cap, il, cl, return_loss=False, return_similarity_mat=True)

                def fwd_bwd():
                    img.grad = cap.grad = None
                    crit(img, cap, il, cl).backward()

                torch.cuda.reset_peak_memory_stats()
                rec = dict(B=B, aggregation=agg, precision=precision, fwd_ms=timeit(fwd, iters),
                           fwd_bwd_ms=timeit(fwd_bwd, iters))
                rec["pairs_per_s_fwd"] = B * B / rec["fwd_ms"] * 1e3
                rec["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
                out.append(rec)
    print(json.dumps({"scan_probe": out}, indent=1))


if __name__ == "__main__":
    main()
