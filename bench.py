#!/usr/bin/env python
"""Benchmark of the all-pairs alignment scoring + Recall@K path (BASELINE.json metric:
"alignment pairs/sec, COCO-5k shape").

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference ...                   # the CPU arm (torch-CPU port of the reference path, oracle/)

One "step" = one full pass of the hot path over the COCO-5k-shape gallery: pack (normalise +
bf16) -> fused tcgen05 MrSw scores for all Ni x Nc pairs -> exact i2t / t2i ranks + top-50 ->
ranks back on the host (the top-50 lists stay on the device until a caller asks for them).  `value` times that with the raw fp32 features resident in HBM;
`e2e` times the public drop-ins (aladin_b200.evaluation.i2t + t2i) on pinned HOST tensors in the
reference layout, H2D copies inside the timed region.  N > 1 shards the gallery images by
contiguous blocks (strong scaling of the fixed 5k problem); rank 0 prints one JSON line."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Ni, Nc, regions, words, d)
    "coco5k": (5000, 25000, 34, 50, 1024),
    "coco1k": (1000, 5000, 34, 50, 1024),
    "tiny": (100, 500, 34, 50, 1024),
}
FLOP_PER_PAIR = lambda regions, words, d: 2.0 * regions * words * d   # noqa: E731  


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ALAD_BENCH_WORKLOAD", "coco5k"), choices=list(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU-baseline budget per sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the `also` block (BASELINE configs 1, 2, 4, 5 and fp32 mode)")
    ap.add_argument("--no-cublas-probe", action="store_true",
                    help="skip the same-box cuBLAS bf16 sustained-GEMM probe reported beside the roofline (context only)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(bf16_burst=d.get("bf16_tflops"), bf16_sustained=d.get("bf16_tflops_sustained"),
                    hbm_gbs=d.get("hbm_gbs"), source="MEASURED_PEAKS.json (measured)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own evaluation loops (alad/evaluation.py:158-327) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_baseline_sample(images_np, captions_np, img_lens, cap_lens, budget_s):
    """Times the reference algorithm on the host cores on a bounded sample: q query images through
    i2t (each against ALL captions, cap_batches=5) and q caption groups through t2i (each against
    ALL images, im_batches=5), all host threads.

    kind "reference": the UNMODIFIED alad/evaluation.py + alad/loss.py copied into oracle/_ref by
    oracle/make_ref.py (run by __graft_entry__.build() where /root/reference exists), driven through
    their public API with the sim_function closure of alad/test.py:259-263 and `.cuda()` as the identity.
    kind "port": oracle/alad_torch_port.py (the same op sequence restated; ~1.3x slower) when the copy is absent.
    Returns (pairs_per_s, description, cores, kind)."""
    import torch
    from oracle import ref_runner as RR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Ni = images_np.shape[0] // 5
    Nc = captions_np.shape[0]
    images_t, captions_t = torch.from_numpy(images_np), torch.from_numpy(captions_np)
    if RR.available():
        kind, what = "reference", "unmodified alad.evaluation.i2t/t2i + alad.loss.AlignmentContrastiveLoss('MrSw') from oracle/_ref"

        def run(q):
            return RR.run_sample(images_t, captions_t, img_lens, cap_lens, q, batches=5)[0]
    else:
        from oracle import alad_torch_port as TP
        kind, what = "port", "torch-CPU port of the reference loops (oracle/alad_torch_port.py; oracle/_ref absent)"

        def run(q):
            t0 = time.perf_counter()
            TP.i2t(images_t, captions_t, img_lens, cap_lens, npts=q, cap_batches=5)
            TP.t2i(images_t, captions_t, img_lens, cap_lens, npts=q, im_batches=5)
            return time.perf_counter() - t0

    t1 = run(1)                                            # also the warm-up
    q = int(max(1, min(Ni, budget_s / max(t1, 1e-3))))
    t = run(q) if q > 1 else t1
    pairs = q * Nc + 5 * q * Ni
    desc = (f"{what}: i2t for {q} query images x {Nc} captions (cap_batches=5) + t2i for {q} caption groups "
            f"({5 * q} captions) x {Ni} images (im_batches=5), fp32, {torch.get_num_threads()} threads, {t:.1f} s")
    return pairs / t, desc, torch.get_num_threads(), kind


def host_layout(images_dev, captions_dev, pinned=True):
    """Reference layout on the host: images [5*Ni, S_im, d] (every image row 5x), captions [Nc, S_s, d]."""
    import torch
    Ni = images_dev.shape[0]
    imgs = torch.empty((5 * Ni,) + tuple(images_dev.shape[1:]), dtype=torch.float32, pin_memory=pinned)
    imgs.view(Ni, 5, *images_dev.shape[1:]).copy_(images_dev.unsqueeze(1).expand(-1, 5, -1, -1))
    caps = torch.empty(tuple(captions_dev.shape), dtype=torch.float32, pin_memory=pinned)
    caps.copy_(captions_dev)
    return imgs, caps


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    Ni, Nc, regions, words, d = WORKLOADS[args.workload]
    config = {"workload": f"{args.workload}: {Ni} images x {Nc} captions all-pairs MrSw alignment scores + i2t/t2i "
                          f"Recall@K, {regions} regions x {words} words, d={d}, dense synthetic features",
              "pairs_per_step": Ni * Nc, "flop_per_pair": FLOP_PER_PAIR(regions, words, d),
              "precision": args.precision, "sharding": f"image blocks over {world} rank(s), captions replicated",
              "l2_policy": "inputs (>= 3 GB per step) are larger than the 126 MB L2; no explicit flush"}

    import torch

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return
        import numpy as np
        from aladin_b200 import synth
        # same seeded generator as our arm, on the CPU here (no GPU needed for this arm)
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d, device=dev)
        imgs_h, caps_h = host_layout(images, captions, pinned=False)
        del images, captions
        img_lens5 = [l for l in im_len for _ in range(5)]
        vals = []
        desc = cores = kind = None
        # every step is a bounded sample; the whole --steps/--warmup run stays within a few minutes
        per_step = min(args.cpu_seconds, 150.0 / max(args.steps + args.warmup, 1))
        for it in range(args.warmup + args.steps):
            budget = per_step if it >= args.warmup else min(per_step, 5.0)
            v, desc, cores, kind = cpu_baseline_sample(imgs_h.numpy(), caps_h.numpy(), img_lens5, s_len, budget)
            if it >= args.warmup:
                vals.append(v)
        value = float(np.mean(vals))
        pairs = Ni * Nc
        line = {"impl": "reference", "metric": "alignment_pairs_per_sec", "value": value, "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * pairs / value,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "gpu_launches": 0,
                "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": desc},
                "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "ms_per_step is the sample rate extrapolated to the full workload (queries are independent)"}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        # NCCL prints its version banner (and any NCCL_DEBUG output) on stdout; keep stdout for the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    group = dist.group.WORLD if world > 1 else None
    import aladin_b200
    from aladin_b200 import _cabi, evaluation, loss as L, retrieval, scoring, synth
    aladin_b200.set_precision(args.precision)
    if os.environ.get("ALAD_NO_POOL"):             # diagnostics: static image blocks, no cross-GPU work pool
        retrieval.POOL = False
    if os.environ.get("ALAD_NO_BALANCE"):          # diagnostics: equal image blocks
        retrieval.balancer.enabled = False
    if os.environ.get("ALAD_DEVICE_PHASES"):       # diagnostics: force the ramped caption phases of the device-resident path on / off
        retrieval.DEVICE_PHASES = os.environ["ALAD_DEVICE_PHASES"] not in ("0", "false", "off")

    images, captions, im_len, s_len = synth.dense_gallery_device(Ni, Nc, regions, words, d)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        _cabi.launch_count["kernels"] = 0
        scoring.kernel_timeline = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        timeline, scoring.kernel_timeline = scoring.kernel_timeline, None
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, timeline, _cabi.launch_count["kernels"], clocks

    # ---- device-resident step: pack + scores + ranks (+ the collectives for N > 1) + ranks to host
    result = {}

    def step_resident():
        gal = retrieval.AlignmentGallery(images, captions, im_len, s_len, n_images=Ni, precision=args.precision,
                                         world=world, rank=rank)
        S = gal.scores(group=group)
        # ranks to the host; top-1 / top-50 are computed (and exchanged between the shards) but stay on the device, as in
        # the public i2t / t2i drop-ins, whose callers read the metrics (alad/test.py:271-276)
        result["out"] = retrieval.rank_both_directions(S, Ni, img_off=gal.lo, n_images_total=Ni, k=50, group=group,
                                                       bounds=gal.bounds, lists_to_host=False)

    ms_step, timeline, launches, clocks = timed(step_resident, args.steps, args.warmup)
    shard_balance = None
    if world > 1:
        sp = retrieval.balancer.speed.get(world)
        pool = next(iter(retrieval._pools.values()), None)
        shard_balance = {"work_pool": ({"tail_fraction": retrieval.POOL_TAIL, "chunks_per_rank": retrieval.POOL_CHUNKS, "chunk_sizes": "decreasing linearly",
                                        "chunks_rank0_own_vs_taken_from_others": dict(pool.stats)} if pool else None),
                         "image_blocks": retrieval.balancer.all_bounds(Ni, world), "relative_speed": sp.tolist() if sp is not None else None,
                         "last_gathered_images_ms": retrieval.balancer.last,
                         "what": "image blocks sized by every rank's measured scoring speed (retrieval.ShardBalancer)"}
    value = Ni * Nc / (ms_step * 1e-3)
    ranks_i2t, _, ranks_t2i, _ = result["out"]
    recalls = {"i2t_r1": retrieval.recall_tuple(ranks_i2t)[0], "t2i_r1": retrieval.recall_tuple(ranks_t2i)[0]}

    # ---- roofline of the dominant kernel (alad::mrsw_fwd_kernel), from events around its launches
    pk = peaks()
    k_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in timeline)
    k_pairs = sum(ni * nc for _, _, ni, nc, _ in timeline)
    k_mult = 3.0 if args.precision == "fp32" else 1.0
    achieved = k_pairs * FLOP_PER_PAIR(regions, words, d) / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    traffic = traffic_detail = None       # DRAM read + write bytes of ONE launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_detail = json.load(f).get(f"{args.workload}_{args.precision}_n{world}")
        if traffic_detail:
            traffic = traffic_detail.get("bytes")
    roofline = {"kernel": "alad::mrsw_fwd_kernel", "bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"],
                "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"] if pk["bf16_sustained"] else None,
                "peak_burst": pk["bf16_burst"], "frac_of_burst": achieved / pk["bf16_burst"] if pk["bf16_burst"] else None,
                "peak_source": pk["source"] + "; sustained figure: the kernel runs ~0.3 s per launch inside the step",
                "launches": len(timeline), "avg_launch_ms": k_ms / max(len(timeline), 1),
                "algorithmic_flop_per_launch": k_pairs * FLOP_PER_PAIR(regions, words, d) / max(len(timeline), 1),
                "issued_flop_multiplier": k_mult, "kernel_share_of_step": k_ms / (ms_step * args.steps), "traffic": traffic,
                "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_detail": traffic_detail}

    # ---- end to end through the public drop-ins, host tensors in the reference layout
    e2e = None
    if not args.no_e2e:
        imgs_h, caps_h = host_layout(images, captions)
        img_lens5 = [l for l in im_len for _ in range(5)]
        scorer = L.AlignmentContrastiveLoss(aggregation="MrSw")
        scorer.precision = args.precision

        def alignment_sim_fn(img, cap, img_len, cap_len):       # what alad/test.py:259-263 builds
            with torch.no_grad():
                return scorer(img, cap, img_len, cap_len, return_loss=False, return_similarity_mat=True)

        def step_e2e():
            evaluation.clear_cache()                            # no reuse across steps
            result["i2t"] = evaluation.i2t(imgs_h, caps_h, img_lens5, s_len, sim_function=alignment_sim_fn, cap_batches=5)
            result["t2i"] = evaluation.t2i(imgs_h, caps_h, img_lens5, s_len, sim_function=alignment_sim_fn, im_batches=5)

        ms_e2e, _, _, _ = timed(step_e2e, max(2, min(args.steps, 3)), 3)
        lo, hi = retrieval.balancer.bounds(Ni, world, rank)
        h2d = (hi - lo) * (regions + 1) * d * 4 + Nc * (words + 1) * d * 4
        d2h = (Ni + Nc) * 8          # both directions' ranks as float64; the top-1 / top-50 lists stay on the device until asked for
        e2e = {"value": Ni * Nc / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "aladin_b200.evaluation.i2t + t2i (sim_function closure over AlignmentContrastiveLoss('MrSw')), "
                      "pinned host tensors [5*Ni,35,1024] / [Nc,53,1024]",
               "recall_at_1": {"i2t": result["i2t"][0], "t2i": result["t2i"][0]}}
        if world > 1:
            e2e["caption_exchange"] = retrieval.EXCHANGE + (" (copy engines through IPC peer windows)" if retrieval.EXCHANGE == "peer" else " all-gather")

    # ---- BASELINE config 5 on the same ranks (N > 1; at N = 1 it is part of the `also` block): two-stage retrieval, K = 100
    two_stage = None
    if world > 1 and not args.no_also and args.workload == "coco5k":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import two_stage_probe
            two_stage = two_stage_probe.measure(Ni, Nc, 100, world=world, steps=3, warmup=2)
        except Exception as e:          # context, never a reason to lose the headline line (all ranks fail alike)
            two_stage = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.no_e2e:
            imgs_h, caps_h = host_layout(images, captions, pinned=False)
            img_lens5 = [l for l in im_len for _ in range(5)]
        v, desc, cores, kind = cpu_baseline_sample(imgs_h.numpy(), caps_h.numpy(), img_lens5, s_len, args.cpu_seconds)
        cpu = {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": desc}

    # ---- context for the roofline: cuBLAS bf16 GEMM sustained on THIS box (after all timed regions)
    if rank == 0 and world == 1 and not args.no_cublas_probe:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import cublas_probe
            probe = cublas_probe.sustained_bf16_tflops(2.5)
            roofline["same_box_cublas_bf16_tflops_sustained"] = probe["cublas_bf16_tflops_sustained"]
            roofline["frac_of_same_box_cublas"] = achieved / probe["cublas_bf16_tflops_sustained"]
        except Exception as e:      # the probe is context, never a reason to lose the bench line
            roofline["same_box_cublas_bf16_tflops_sustained"] = f"probe failed: {e}"

    # ---- e2e once more from PAGEABLE host tensors (what the reference's encode_data hands over)
    if e2e is not None and rank == 0 and world == 1 and not args.no_also:
        imgs_p, caps_p = host_layout(images, captions, pinned=False)

        def step_pageable():
            evaluation.clear_cache()
            evaluation.i2t(imgs_p, caps_p, img_lens5, s_len, sim_function=alignment_sim_fn, cap_batches=5)
            evaluation.t2i(imgs_p, caps_p, img_lens5, s_len, sim_function=alignment_sim_fn, im_batches=5)

        ms_pg, _, _, _ = timed(step_pageable, 2, 1)
        e2e["pageable_host_tensors"] = {"ms_per_step": ms_pg, "value": Ni * Nc / (ms_pg * 1e-3), "unit": "pairs/s"}
        del imgs_p, caps_p

    # ---- the other BASELINE configs in the same run (N = 1 only)
    also = None
    if rank == 0 and world == 1 and not args.no_also:
        del images, captions
        torch.cuda.empty_cache()
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_also
            also = bench_also.also_block(pk)
        except Exception as e:          # context, never a reason to lose the headline line
            also = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {"metric": "alignment_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "cpu_baseline": cpu, "recall_at_1": recalls, "also": also, "two_stage": two_stage if world > 1 else (also or {}).get("config5_two_stage_coco5k"),
                "shard_balance": shard_balance,
                "tflops_algorithmic_whole_step": Ni * Nc * FLOP_PER_PAIR(regions, words, d) / (ms_step * 1e-3) / 1e12}
        print(json.dumps(line))
    if world > 1:
        retrieval.close_exchanges()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
