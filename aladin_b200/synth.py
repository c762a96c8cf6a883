"""Seeded synthetic inputs of the shapes SURVEY.md §8(d) / BASELINE.json name.  Used by the
tests and by bench.py; CPU-generated (numpy RandomState) for small cases so that the oracle
and the CUDA path see identical bits, torch-generated on the device for the full-size sets."""
import numpy as np


def raw_batch(seed, Bi, Bc, S_im, S_s, d, ragged=True, related=0.0):
    """Training-shaped raw tensors [B,S,d] fp32 + python length lists (reference layout:
    image slot 0 and caption slots 0, -2, -1 are never scored: alad/loss.py:87-90)."""
    r = np.random.RandomState(seed)
    im = r.standard_normal((Bi, S_im, d)).astype(np.float32)
    s = r.standard_normal((Bc, S_s, d)).astype(np.float32)
    if related:
        n = min(Bi, Bc)
        k = min(S_im, S_s) - 1
        s[:n, 1:k] += related * im[:n, 1:k]
    if ragged:
        im_len = r.randint(2, S_im + 1, size=Bi)
        s_len = r.randint(4, S_s + 1, size=Bc)
        im_len[r.randint(Bi)] = S_im           # at least one full-length item (no clamp)
        s_len[r.randint(Bc)] = S_s
    else:
        im_len = np.full(Bi, S_im)
        s_len = np.full(Bc, S_s)
    return im, s, [int(x) for x in im_len], [int(x) for x in s_len]


def eval_containers(seed, Ni, S, d, max_regions, max_words, alpha=0.55, dense=False):
    """Evaluation containers like encode_data (alad/evaluation.py:98-130): [5*Ni,S,d] fp32,
    slot 0 = global vector, tokens from slot 1, zero padding, each image row repeated 5x.
    Caption c is noise + alpha * (random regions of image c//5) so recalls are non-trivial."""
    r = np.random.RandomState(seed)
    N = 5 * Ni
    if dense:
        img_feat_len = np.full(Ni, max_regions)
        cap_len = np.full(N, max_words)
    else:
        img_feat_len = r.randint(3, max_regions + 1, size=Ni)
        cap_len = r.randint(4, max_words + 1, size=N)
    base = r.standard_normal((Ni, S, d)).astype(np.float32)
    captions = np.zeros((N, S, d), np.float32)
    for i in range(Ni):
        base[i, img_feat_len[i]:] = 0
    for c in range(N):
        L = int(cap_len[c])
        noise = r.standard_normal((L, d)).astype(np.float32)
        src = base[c // 5, 1 + r.randint(0, max(img_feat_len[c // 5] - 1, 1), size=L)]
        captions[c, :L] = noise + alpha * src
    images = np.repeat(base, 5, axis=0)
    img_lens = [int(img_feat_len[i // 5]) for i in range(N)]
    return images, captions, img_lens, [int(x) for x in cap_len]


def dense_gallery_device(Ni, Nc, regions=34, words=50, d=1024, device="cuda", alpha=0.05, seeds=(1234, 5678)):
    """Full-size dense roofline set generated on the device: every image has `regions`
    scored regions, every caption `words` scored words -> raw containers S_im = regions+1,
    S_s = words+3.  Returns (images[Ni,S_im,d], captions[Nc,S_s,d], im_len, s_len);
    image i here is gallery image i (i.e. row 5i of the reference's 5x-duplicated tensor)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seeds[0])
    images = torch.randn((Ni, regions + 1, d), generator=g, device=device, dtype=torch.float32)
    images = torch.nn.functional.normalize(images, dim=2)
    g.manual_seed(seeds[1])
    captions = torch.empty((Nc, words + 3, d), device=device, dtype=torch.float32)
    step = 1000
    group = max(Nc // max(Ni, 1), 1)
    for c0 in range(0, Nc, step):
        c1 = min(Nc, c0 + step)
        noise = torch.randn((c1 - c0, words + 3, d), generator=g, device=device, dtype=torch.float32)
        owner = (torch.arange(c0, c1, device=device) // group).clamp_(max=Ni - 1)
        pick = torch.randint(1, regions + 1, (c1 - c0, words + 3), generator=g, device=device)
        src = images[owner.unsqueeze(1), pick]                       # [n, S_s, d]
        captions[c0:c1] = torch.nn.functional.normalize(noise / (d ** 0.5) + alpha * src, dim=2)
    return images, captions, [regions + 1] * Ni, [words + 3] * Nc
