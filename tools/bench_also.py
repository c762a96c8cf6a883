"""The `also` block of bench.py's JSON line: every BASELINE.json config other than the headline one, measured in the
same run on the same box (N = 1, rank 0), so that the driver sees them:

  config 1  training-batch losses, B = 128 (fwd, fwd+bwd)          -- fused forward_loss path (alad_model.train_losses)
  config 2  COCO-1k retrieval step (1000 x 5000)                   -- same step as the headline, smaller gallery
  config 3  (headline) in fp32 mode = bf16 x 3 split precision      -- own roofline fraction (issued FLOP x 3)
  config 4  training step at B = 512 (fwd, fwd+bwd)                 -- with the tensor-roofline fraction of the forward
  config 5  two-stage retrieval at COCO-5k shape, K = 100           -- pair-list stage 2 (tools/two_stage_probe.py)

All timings: CUDA events on the launching stream after warm-up, synchronised on both sides."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def retrieval_step(Ni, Nc, precision, steps=3, warmup=3, regions=34, words=50, d=1024):
    """pack + scores + ranks of one gallery with device-resident raw features: (ms per step, kernel ms per step)."""
    from aladin_b200 import retrieval, scoring, synth
    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d)

    def step():
        gal = retrieval.AlignmentGallery(images, captions, im_len, s_len, n_images=Ni, precision=precision)
        S = gal.scores()
        retrieval.rank_both_directions(S, Ni, k=50)

    for _ in range(warmup):
        step()
    scoring.kernel_timeline = []
    ms = _timed(step, steps, 0)
    tl, scoring.kernel_timeline = scoring.kernel_timeline, None
    k_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in tl) / steps
    del images, captions
    return ms, k_ms


def also_block(pk, quick=False):
    import bench_train_step as TS
    import two_stage_probe as P2
    flop_pair = 2.0 * 34 * 50 * 1024
    out = {}
    # ---- configs 1 and 4: the three-criterion training step (alad/alad_model.py:371-428)
    for name, B in (("config1_train_B128", 128), ("config4_train_B512", 512)):
        fwd, fwd_bwd, il, cl = TS.make_step(B, "bf16", fused=True)
        with torch.no_grad():
            t_f = TS.timeit(fwd, iters=50, warm=10)
        t_fb = TS.timeit(fwd_bwd, iters=50, warm=10)
        flop = 2.0 * sum(l - 1 for l in il) * sum(l - 3 for l in cl) * 1024
        out[name] = {"fwd_ms": t_f, "fwd_bwd_ms": t_fb, "pairs_per_s_fwd": B * B / t_f * 1e3,
                     "pairs_per_s_fwd_bwd": B * B / t_fb * 1e3, "fwd_algorithmic_tflops": flop / t_f / 1e9,
                     "fwd_frac_of_bf16_burst_peak": flop / t_f / 1e9 / pk["bf16_burst"] if pk.get("bf16_burst") else None,
                     "what": f"matching + MrSw alignment + hinge + ListNet, B={B}, ragged 35x53 slots, d=1024, bf16, "
                             "fused forward_loss path (one native call per direction)"}
    # ---- config 4 with max_violation=False (alad/loss.py:60-67): dL/dS is dense, the sparse pair-list backward has to
    # recompute all B^2 pairs on CUDA cores (missing 5: no tensor-core dense-G backward)
    fwd, fwd_bwd, il, cl = TS.make_step(512, "bf16", fused=True, max_violation=False)
    with torch.no_grad():
        t_f = TS.timeit(fwd, iters=20, warm=5)
    t_fb = TS.timeit(fwd_bwd, iters=10, warm=3)
    out["config4_train_B512_sum_of_violations"] = {"fwd_ms": t_f, "fwd_bwd_ms": t_fb,
                                                   "what": "same step with max_violation=False: dense dL/dS backward"}
    import aladin_b200
    aladin_b200.set_precision("bf16")
    # ---- config 2: COCO-1k shape retrieval step
    ms, k_ms = retrieval_step(1000, 5000, "bf16", steps=10, warmup=3)
    out["config2_coco1k"] = {"ms_per_step": ms, "pairs_per_s": 5e6 / (ms * 1e-3), "kernel_ms": k_ms,
                             "kernel_tflops": 5e6 * flop_pair / (k_ms * 1e-3) / 1e12,
                             "frac_of_bf16_burst_peak": 5e6 * flop_pair / (k_ms * 1e-3) / 1e12 / pk["bf16_burst"]}
    # ---- config 3 in fp32 mode: three bf16 products per dot product (hi*hi + hi*lo + lo*hi)
    if not quick:
        ms, k_ms = retrieval_step(5000, 25000, "fp32", steps=2, warmup=2)
        algo = 1.25e8 * flop_pair / (k_ms * 1e-3) / 1e12
        out["config3_coco5k_fp32_mode"] = {
            "ms_per_step": ms, "pairs_per_s": 1.25e8 / (ms * 1e-3), "kernel_ms": k_ms, "algorithmic_tflops": algo,
            "issued_tflops": 3 * algo, "frac_issued_of_bf16_sustained_peak": 3 * algo / pk["bf16_sustained"],
            "frac_algorithmic_of_bf16_sustained_peak": algo / pk["bf16_sustained"],
            "what": "fp32-grade scores (<= 1e-4 rel) from split-precision bf16 operands: 3x the tensor FLOP of bf16 mode"}
        # the same step on the TF32 tensor path (fp32 operands rounded to TF32, kind::tf32): an intermediate precision (worst
        # entry 3e-4 .. 6e-4 relative; the 1e-4 parity mode is the split-precision one above)
        ms, k_ms = retrieval_step(5000, 25000, "tf32", steps=2, warmup=2)
        algo = 1.25e8 * flop_pair / (k_ms * 1e-3) / 1e12
        out["config3_coco5k_tf32_mode"] = {
            "ms_per_step": ms, "pairs_per_s": 1.25e8 / (ms * 1e-3), "kernel_ms": k_ms, "algorithmic_tflops": algo,
            "frac_of_half_the_bf16_sustained_peak": algo / (0.5 * pk["bf16_sustained"]),
            "what": "fp32 operands read as TF32 by the tensor core (half the bf16 rate, twice the operand bytes)"}
    # ---- config 5: two-stage retrieval
    out["config5_two_stage_coco5k"] = P2.measure(5000, 25000, 100, world=1, steps=3, warmup=2)
    # ---- the HBM-bound kernels around the scoring kernel (north_star: "achieved HBM GB/s for the loss and top-k kernels"):
    # ranking of a COCO-5k-sized score matrix (two reads of S) and the B x B losses, CUDA events, algorithmic bytes.
    # Own guard: a failure here must not take the configs above with it.
    try:
        import hbm_probe as HP
        rk = HP.ranking_part()
        ls = [HP.loss_part(8192), HP.loss_part(512)]
        hbm = pk.get("hbm_gbs")
        out["hbm_kernels"] = {
            "ranking_coco5k": {"ms": rk["fused_ms"], "reads_of_S": 2, "one_purpose_kernels_ms": rk["separate_ms"],
                               "algorithmic_bytes": rk["sweep_bytes"], "GBs_algorithmic": rk["fused_GBs_algorithmic"],
                               "frac_of_hbm_peak": rk["fused_GBs_algorithmic"] / hbm if hbm else None,
                               "equals_one_purpose_kernels": rk["fused_equals_separate"] and rk["masked_equal"]},
            "losses": [{"B": l["B"], "triplet_ms": l["triplet_ms"], "triplet_GBs_algorithmic": l["triplet_GBs_algorithmic"],
                        "listnet_ms": l["listnet_ms"], "listnet_GBs_algorithmic": l["listnet_GBs_algorithmic"],
                        "listnet_frac_of_hbm_peak": l["listnet_GBs_algorithmic"] / hbm if hbm else None,
                        "listnet_loss_rel_err_vs_fp64": l["listnet_loss_rel_err"],
                        "listnet_grad_err_vs_fp64": l["listnet_grad_max_err_rel_to_max"]} for l in ls],
            "what": "CUDA-event times; algorithmic bytes = 4*Ni*Nc (one reading of S), 8*B^2 (triplet), 16*B^2 (listnet)"}
    except Exception as e:                          # noqa: BLE001
        out["hbm_kernels"] = {"error": f"{type(e).__name__}: {e}"}
    return out
