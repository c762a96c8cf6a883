"""Per-phase CUDA-event timeline of the end-to-end multi-rank step (host-resident captions: upload + pack + exchange of
the packed rows in column phases, overlapped with scoring).  Run under torchrun; every rank prints one JSON line with,
per phase, the offsets (ms from the step's start) of: prep start, packed, gathered, scoring start, scoring end.
    python -m torch.distributed.run --nproc-per-node N tools/e2e_timeline.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from aladin_b200 import evaluation, loss as L, retrieval, synth
    Ni, Nc, regions, words, d = bench.WORKLOADS[os.environ.get("ALAD_BENCH_WORKLOAD", "coco5k")]
    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d)
    imgs_h, caps_h = bench.host_layout(images, captions)
    group = dist.group.WORLD

    # ---- the device-resident step of bench.py (`value`): pack + scores + ranking exchange + results to the host
    def step_resident():
        gal = retrieval.AlignmentGallery(images, captions, im_len, s_len, n_images=Ni, precision="bf16", world=world, rank=rank)
        S = gal.scores(group=group)
        return retrieval.rank_both_directions(S, Ni, img_off=gal.lo, n_images_total=Ni, k=50, group=group, bounds=gal.bounds,
                                              lists_to_host=False)

    from aladin_b200 import scoring
    for _ in range(4):
        step_resident()
    dist.barrier()
    torch.cuda.synchronize()
    res_rows = []
    for _ in range(3):
        retrieval.rank_timeline, scoring.kernel_timeline = [], []
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        step_resident()
        r1.record()
        torch.cuda.synchronize()
        marks = retrieval.rank_timeline[0]
        k0, k1 = scoring.kernel_timeline[0][0], scoring.kernel_timeline[-1][1]
        res_rows.append({"step_ms": round(r0.elapsed_time(r1), 3), "kernel_start": round(r0.elapsed_time(k0), 3),
                         "kernel_end": round(r0.elapsed_time(k1), 3), **{k: round(r0.elapsed_time(ev), 3) for k, ev in marks.items()}})
    retrieval.rank_timeline, scoring.kernel_timeline = None, None
    del images, captions
    img_lens5 = [l for l in im_len for _ in range(5)]
    scorer = L.AlignmentContrastiveLoss(aggregation="MrSw")
    fn = evaluation.fused_sim_function(scorer)
    if os.environ.get("ALAD_EXCHANGE"):
        retrieval.EXCHANGE = os.environ["ALAD_EXCHANGE"]

    def step():
        evaluation.clear_cache()
        evaluation.i2t(imgs_h, caps_h, img_lens5, s_len, sim_function=fn, cap_batches=5)
        evaluation.t2i(imgs_h, caps_h, img_lens5, s_len, sim_function=fn, im_batches=5)

    for _ in range(4):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    retrieval.phase_timeline = []
    retrieval.rank_timeline = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    tl, retrieval.phase_timeline = retrieval.phase_timeline, None
    rtl, retrieval.rank_timeline = retrieval.rank_timeline, None
    ranking = [{k: round(e0.elapsed_time(ev), 3) for k, ev in m.items()} for m in rtl]
    rows = [{"captions": [p["c0"], p["c1"]], **{k: round(e0.elapsed_time(p[k]), 3) for k in ("t0", "packed", "gathered", "s0", "s1")}}
            for p in tl]
    out = {"rank": rank, "world": world, "exchange": retrieval.EXCHANGE, "step_ms": round(e0.elapsed_time(e1), 3), "phases": rows, "ranking": ranking, "resident_steps": res_rows}
    for r in range(world):
        if r == rank:
            print(json.dumps(out), flush=True)
        dist.barrier()
    retrieval.close_exchanges()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
