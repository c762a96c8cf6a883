"""Two-stage retrieval (config 5) at COCO-5k shape on one GPU (or under torchrun): time of the whole call, of stage 2's
pair-list kernel, the number of pair tiles, against the dense pass.  Prints one JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from aladin_b200 import scoring, synth, two_stage
    Ni, Nc = int(os.environ.get("NI", 5000)), int(os.environ.get("NC", 25000))
    res = measure(Ni, Nc, 100, world)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def measure(Ni, Nc, K, world=1, steps=5, warmup=2, regions=34, words=50, d=1024):
    """Synthetic config-5 gallery on the device: dense COCO-shape tokens; the slot-0 global vector of an image is the
    normalised mean of its regions and that of a caption the normalised mean of its words, so that the matching head
    ranks like a (weaker) alignment head, as in the trained model."""
    import torch.distributed as dist
    from aladin_b200 import scoring, synth, two_stage
    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d)
    images[:, 0, :] = torch.nn.functional.normalize(images[:, 1:, :].mean(dim=1), dim=1)
    captions[:, 0, :] = torch.nn.functional.normalize(captions[:, 1:1 + words, :].mean(dim=1), dim=1)
    imgs5 = images.unsqueeze(1).expand(-1, 5, -1, -1).reshape(5 * Ni, regions + 1, d)   # a view-free 5x layout is not needed:
    img_lens5 = [l for l in im_len for _ in range(5)]                                    # rows 0::5 are what is read
    out = {}

    def step():
        out["res"] = two_stage.two_stage_retrieval(imgs5, captions, img_lens5, s_len, shortlist=K, precision="bf16",
                                                   return_details=True)

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    scoring.kernel_timeline = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tl, scoring.kernel_timeline = scoring.kernel_timeline, None
    pair_ms = [a.elapsed_time(b) for a, b, n, _, _ in tl if isinstance(n, torch.Tensor)]
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # one more call with stage marks
    two_stage.stage_timeline = []
    step()
    torch.cuda.synchronize()
    tl, two_stage.stage_timeline = two_stage.stage_timeline, None
    stages = {b[0]: round(a[1].elapsed_time(b[1]), 3) for a, b in zip(tl[:-1], tl[1:])}
    (m_i2t, m_t2i), det = out["res"]
    n_tiles = int(det["n_ptiles"].item()) if det["n_ptiles"] is not None else 0
    pairs = Ni * min(K, Nc) + Nc * min(K, Ni)
    slots = 240 // two_stage.slot_rows_for(torch.tensor([regions]).numpy())
    flop_issued = n_tiles * 2.0 * 128 * 240 * d
    k_ms = sum(pair_ms) / max(len(pair_ms), 1)
    return {"workload": f"two-stage {Ni}x{Nc}, K={K}, {regions}x{words} tokens, d={d}, bf16", "n_gpus": world,
            "ms_per_call": ms, "shortlisted_pairs": pairs, "pairs_per_s": pairs / (ms * 1e-3),
            "pair_kernel_ms": k_ms, "pair_tiles_rank0": n_tiles, "slots_per_tile": slots,
            "pair_kernel_issued_tflops": flop_issued / (k_ms * 1e-3) / 1e12 if k_ms else None,
            "algorithmic_flop": pairs * 2.0 * regions * words * d, "stage_ms": stages,
            "recall_at_1": {"i2t": m_i2t[0], "t2i": m_t2i[0]}}


if __name__ == "__main__":
    main()
