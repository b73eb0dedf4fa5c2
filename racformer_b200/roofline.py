"""Algorithmic-byte model of the sampling kernels (SURVEY.md section 8d reference card; stated in DESIGN.md).

The unit of work is a tap (one sample point on one level). Bytes are counted from the validity masks of the actual
inputs (`racf_msmv_tap_masks` / `racf_msda_tap_masks`: bit0 = in range, bits 1..4 = corner read), so the figure is
algorithmic traffic, not measured DRAM traffic.
"""
import torch


def _corner_count(mask):
    m = mask.to(torch.int32)
    return int((((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1) + ((m >> 4) & 1)).sum())


def msmv_bytes(tap_mask, C, L, feat_bytes):
    """tap_mask uint8 [B,Q,P,L]. Returns (forward_bytes, backward_bytes) per call.

    forward : every read corner is C*4 bytes; per sample point loc 12 B + weights 4L B read, C*4 B written.
    backward: per point grad_out C*4 + loc/weights (12+4L) read and grad loc/weights (12+4L) written; per read
              corner C*4 (feature re-read) + 2*C*4 (read-modify-write of the feature gradient); plus one zero-fill
              of all feature-gradient maps (feat_bytes).
    """
    corners = _corner_count(tap_mask)
    npts = tap_mask.numel() // L
    fwd = corners * C * 4 + npts * (12 + 4 * L + C * 4)
    bwd = npts * (C * 4 + 2 * (12 + 4 * L)) + corners * C * 4 * 3 + feat_bytes
    return fwd, bwd


def msda_bytes(tap_mask, D, value_bytes):
    """tap_mask uint8 [B,Q,M,L,P]. Returns (forward_bytes, backward_bytes) per call.

    forward : read corners D*4 each; per tap loc 8 B + weight 4 B; per (b,q,head) D*4 written.
    backward: per read corner D*4 + 2*D*4; per tap 12 B read + 12 B written; per (b,q,head) D*4 grad_out read;
              plus the zero-fill of grad_value (value_bytes) that the caller performs.
    """
    corners = _corner_count(tap_mask)
    ntaps = tap_mask.numel()
    nrows = tap_mask.shape[0] * tap_mask.shape[1] * tap_mask.shape[2]
    fwd = corners * D * 4 + ntaps * 12 + nrows * D * 4
    bwd = corners * D * 4 * 3 + ntaps * 24 + nrows * D * 4 + value_bytes
    return fwd, bwd
