"""Shared helpers for the test-suite: golden fixtures, seeded input generators, tolerant comparison."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MSMV_CASES = ["msmv_l4_c8", "msmv_l4_c64", "msmv_l2_c64", "msmv_l5_c64"]
MSDA_CASES = ["msda_l1_d64", "msda_l3_d64", "msda_l2_d32"]

# Tolerances (measured on the reference itself): see DESIGN.md "Parity".
FWD_RTOL = 1e-5
FWD_ATOL_EXACT = 1e-6      # x max|feat| : CUDA vs exact-arithmetic C oracle
FWD_ATOL_GRIDSAMPLE = 5e-5  # x max|feat| : anything vs the reference's grid_sample fallback
BWD_RTOL = 1e-4
BWD_ATOL = 1e-5            # x max|expected| of the tensor compared (atomic / shuffle-tree reordering)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def msmv_feats(d):
    n = sum(1 for k in d if k.startswith("feat"))
    return [d[f"feat{i}"] for i in range(n)]


def assert_close(actual, expected, rtol, atol, what):
    actual, expected = actual.detach().cpu().double(), expected.detach().cpu().double()
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} vs {tuple(expected.shape)}"
    err = (actual - expected).abs()
    bound = atol + rtol * expected.abs()
    bad = err > bound
    if bad.any():
        i = int(torch.argmax(err - bound))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance (rtol={rtol}, atol={atol:.3g}); "
            f"worst |err|={float(err.flatten()[i]):.3g} at expected={float(expected.flatten()[i]):.6g}")


def near_integer_pixel_msmv(loc, shapes, eps=1e-3):
    """Points whose pixel coordinate is within eps of an integer on some level: d/dloc is one-sided there and
    grid_sample's `(2x-1+1)/2*(W-1)` rounding may pick the other side than the kernel's `x*(W-1)`."""
    bad = torch.zeros(loc.shape[:-1], dtype=torch.bool)
    for h, w in shapes:
        for coord, size in ((loc[..., 0], w), (loc[..., 1], h)):
            if size == 1:   # d(pixel)/d(loc) = size-1 = 0: this axis never contributes to grad_loc
                continue
            pix = coord.double() * (size - 1)
            bad |= (pix - pix.round()).abs() < eps
    return bad


def near_integer_pixel_msda(loc, shapes, eps=1e-3):
    bad = torch.zeros(loc.shape[:-1], dtype=torch.bool)
    for l, (h, w) in enumerate(shapes):
        for k, size in ((0, w), (1, h)):
            pix = loc[:, :, :, l, :, k].double() * size - 0.5
            bad[:, :, :, l, :] |= (pix - pix.round()).abs() < eps
    return bad


def make_msmv_inputs(seed, Bp, N, C, Q, P, shapes, lo=0.0, hi=1.0, device="cpu"):
    """Seeded synthetic MSMV inputs in the CUDA layout."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(Bp, N, h, w, C, generator=g) for h, w in shapes]
    xy = torch.rand(Bp, Q, P, 2, generator=g) * (hi - lo) + lo
    view = torch.randint(0, N, (Bp, Q, P, 1), generator=g).float() / (N - 1)
    loc = torch.cat([xy, view], -1).contiguous()
    w = torch.softmax(torch.randn(Bp, Q, P, len(shapes), generator=g), -1).contiguous()
    grad_out = torch.randn(Bp, Q, C, P, generator=g)
    mv = lambda t: t.to(device)
    return [mv(f) for f in feats], mv(loc), mv(w), mv(grad_out)


def make_msda_inputs(seed, B, M, D, Q, P, shapes, lo=0.0, hi=1.0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(B, S, M, D, generator=g)
    loc = (torch.rand(B, Q, M, L, P, 2, generator=g) * (hi - lo) + lo).contiguous()
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), -1).view(B, Q, M, L, P).contiguous()
    grad_out = torch.randn(B, Q, M * D, generator=g)
    spatial = torch.tensor(shapes, dtype=torch.int64)
    starts = [0]
    for h, w in shapes[:-1]:
        starts.append(starts[-1] + h * w)
    lsi = torch.tensor(starts, dtype=torch.int64)
    mv = lambda t: t.to(device)
    return mv(value), mv(spatial), mv(lsi), mv(loc), mv(aw), mv(grad_out)


F8_SHAPES = [(64, 176), (32, 88), (16, 44), (8, 22)]
