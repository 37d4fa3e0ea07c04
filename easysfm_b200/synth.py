"""Seeded synthetic descriptor banks of the BASELINE.json shapes.

SURF-like: a world of `world_factor * F` landmarks ~ N(0, I_64), L2-normalised; every image samples F
landmarks without replacement, adds per-dimension N(0, sigma^2) noise with sigma drawn per image from
`sigmas`, and re-normalises -> true matches at d ~ 0.08-0.25, exactly where the fp32 expansion form loses
relative precision.
ORB-like: a world of 4F uniform 256-bit strings; every image samples F and flips each bit with p = 0.08.
Uniform bits concentrate impostor distances at 128 +- 8, so rank-2 ties are everywhere -- the lowest-index
tie-break is exercised on almost every query.
"""
from __future__ import annotations

import numpy as np


def surf_like(n_images: int, n_feat, seed: int = 2, sigmas=(0.01, 0.03), world_factor: int = 4):
    """-> list of float32 arrays [F_i, 64].  `n_feat` may be an int or a per-image sequence (ragged)."""
    rng = np.random.default_rng(seed)
    feats = [int(n_feat)] * n_images if np.isscalar(n_feat) else [int(x) for x in n_feat]
    fmax = max(feats) if feats else 0
    W = max(world_factor * fmax, 2)
    world = rng.standard_normal((W, 64), dtype=np.float32)
    world /= np.linalg.norm(world, axis=1, keepdims=True)
    out = []
    for f in feats:
        ids = rng.choice(W, size=f, replace=False)
        sigma = sigmas[int(rng.integers(len(sigmas)))]
        d = world[ids] + sigma * rng.standard_normal((f, 64), dtype=np.float32)
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
        out.append(np.ascontiguousarray(d.astype(np.float32)))
    return out


def orb_like(n_images: int, n_feat, seed: int = 3, flip_p: float = 0.08, world_factor: int = 4):
    """-> list of uint8 arrays [F_i, 32]."""
    rng = np.random.default_rng(seed)
    feats = [int(n_feat)] * n_images if np.isscalar(n_feat) else [int(x) for x in n_feat]
    fmax = max(feats) if feats else 0
    W = max(world_factor * fmax, 2)
    world = rng.integers(0, 256, size=(W, 32), dtype=np.uint8)
    out = []
    for f in feats:
        ids = rng.choice(W, size=f, replace=False)
        flips = np.packbits(rng.random((f, 256)) < flip_p, axis=1)
        out.append(np.ascontiguousarray(world[ids] ^ flips))
    return out


def surf_like_torch(n_images: int, n_feat: int, seed: int, device, sigma: float = 0.02, world_factor: int = 4):
    """Same family, generated on the GPU with torch's Philox stream (bench-sized banks: 1000 x 8000 x 64 in < 1 s).
    -> float32 tensor [n_images, n_feat, 64] on `device`."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    W = world_factor * n_feat
    world = torch.randn((W, 64), generator=g, device=device, dtype=torch.float32)
    world /= world.norm(dim=1, keepdim=True)
    out = torch.empty((n_images, n_feat, 64), device=device, dtype=torch.float32)
    for i in range(n_images):
        ids = torch.randperm(W, generator=g, device=device)[:n_feat]
        d = world[ids] + sigma * torch.randn((n_feat, 64), generator=g, device=device, dtype=torch.float32)
        out[i] = d / d.norm(dim=1, keepdim=True)
    return out


def orb_like_torch(n_images: int, n_feat: int, seed: int, device, flip_p: float = 0.08, world_factor: int = 4):
    """-> uint8 tensor [n_images, n_feat, 32] on `device`."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    W = world_factor * n_feat
    world = torch.randint(0, 256, (W, 32), generator=g, device=device, dtype=torch.uint8)
    weights = (2 ** torch.arange(7, -1, -1, device=device)).to(torch.int32)
    out = torch.empty((n_images, n_feat, 32), device=device, dtype=torch.uint8)
    for i in range(n_images):
        ids = torch.randperm(W, generator=g, device=device)[:n_feat]
        bits = (torch.rand((n_feat, 32, 8), generator=g, device=device) < flip_p).to(torch.int32)
        flips = (bits * weights).sum(dim=2).to(torch.uint8)
        out[i] = world[ids] ^ flips
    return out
