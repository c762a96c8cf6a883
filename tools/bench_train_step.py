"""Training-side measurements (BASELINE configs 1 and 4): alignment + matching + hinge + ListNet,
forward and forward+backward, through the drop-in criteria; plus achieved HBM GB/s of the B x B
loss kernels at B = 512 and B = 8192 (algorithmic bytes: triplet 8*B^2, listnet 16*B^2).
Run on the GPU box:  python tools/bench_train_step.py"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import aladin_b200  # noqa: E402
from aladin_b200 import alad_model as AM, loss as L, synth  # noqa: E402


def timeit(fn, iters=100, warm=10):
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


def make_step(B, precision, fused=False, max_violation=True):
    """(fwd, fwd_bwd, il, cl): the three-criterion training step of alad_model.py:371-428 on synthetic features."""
    im, s, il, cl = synth.raw_batch(9, B, B, 35, 53, 1024, related=0.6)
    r = np.random.RandomState(1)
    icls = r.standard_normal((B, 1024)).astype(np.float32)
    ccls = (0.7 * icls + r.standard_normal((B, 1024))).astype(np.float32)
    icls /= np.linalg.norm(icls, axis=1, keepdims=True)
    ccls /= np.linalg.norm(ccls, axis=1, keepdims=True)
    img_set = torch.tensor(im.transpose(1, 0, 2).copy(), device="cuda", requires_grad=True)   # [S,B,d]
    cap_seq = torch.tensor(s.transpose(1, 0, 2).copy(), device="cuda", requires_grad=True)
    img_cls = torch.tensor(icls, device="cuda", requires_grad=True)
    cap_cls = torch.tensor(ccls, device="cuda", requires_grad=True)
    mc = L.ContrastiveLoss(margin=0.2, measure="dot", max_violation=max_violation)
    ac = L.AlignmentContrastiveLoss(margin=0.2, measure="dot", max_violation=max_violation, aggregation="MrSw")
    dl = L.DistillationLoss(mode="listnet")
    aladin_b200.set_precision(precision)

    def fwd():
        if fused:       # aladin_b200.alad_model: one native call per direction (the forward_loss call site)
            ml, al, d, _, _ = AM.train_losses(img_cls, cap_cls, img_set.permute(1, 0, 2), cap_seq.permute(1, 0, 2), il, cl,
                                              margin=0.2, max_violation=max_violation)
            return al + d + 0.1 * ml
        ml, mm = mc(img_cls, cap_cls, return_similarity_mat=True)
        al, ts = ac(img_set.permute(1, 0, 2), cap_seq.permute(1, 0, 2), il, cl, return_similarity_mat=True)
        return al + dl(ts, mm) + 0.1 * ml

    def fwd_bwd():
        for t in (img_set, cap_seq, img_cls, cap_cls):
            t.grad = None
        fwd().backward()

    return fwd, fwd_bwd, il, cl


def train_step(B, precision, fused=False):
    fwd, fwd_bwd, il, cl = make_step(B, precision, fused)
    with torch.no_grad():
        t_f = timeit(fwd)
    t_fb = timeit(fwd_bwd)
    # host enqueue time per step: wall clock of the Python side with an empty launch queue
    import time
    host = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fwd_bwd()
        host.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    t_host = 1e3 * sorted(host)[len(host) // 2]
    flop = 2.0 * sum(l - 1 for l in il) * sum(l - 3 for l in cl) * 1024
    return {"B": B, "precision": precision, "path": "fused forward_loss" if fused else "per-criterion drop-ins", "fwd_ms": t_f, "fwd_bwd_ms": t_fb, "fwd_bwd_host_enqueue_ms": t_host, "pairs_per_s_fwd": B * B / t_f * 1e3,
            "pairs_per_s_fwd_bwd": B * B / t_fb * 1e3, "fwd_algorithmic_tflops": flop / t_f / 1e9}


def loss_kernels(B):
    r = np.random.RandomState(B)
    S = torch.tensor(r.standard_normal((B, B)).astype(np.float32), device="cuda")
    M = torch.tensor(np.clip(r.standard_normal((B, B)) * 0.3, -1, 1).astype(np.float32), device="cuda")
    T = S * 2 + 3
    t_tri = timeit(lambda: L.triplet_fwd_bwd(S, 0.2, True))
    t_ln = timeit(lambda: L.listnet_fwd_bwd(T, M))
    return {"B": B, "triplet_ms": t_tri, "triplet_GBs": 8.0 * B * B / t_tri / 1e6, "listnet_ms": t_ln,
            "listnet_GBs": 16.0 * B * B / t_ln / 1e6}


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "profile":      # ncu launch list: 1 warm-up + 1 profiled step
        _, step, _, _ = make_step(int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "bf16", fused=len(sys.argv) > 4)
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "hostprof":     # where the host time of a step goes (cProfile, no sync inside)
        import cProfile
        import pstats
        _, step, _, _ = make_step(int(sys.argv[2]), "bf16", fused=len(sys.argv) > 3)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(50):
            step()
            torch.cuda.synchronize()        # empty launch queue: host time is not back-pressure from the device
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
        sys.exit(0)
    out = {"train_step": [train_step(B, p, f) for B in (128, 512) for p in ("bf16", "fp32") for f in (False, True)],
           "loss_kernels": [loss_kernels(B) for B in (512, 8192)]}
    print(json.dumps(out, indent=1))
