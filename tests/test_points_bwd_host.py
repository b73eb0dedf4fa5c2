"""CPU: the hand-derived backward rules of the fused sampling-point chain (racformer_b200/csrc/points_bwd.cuh -- the
functions the CUDA kernels of csrc/points_train.cu call) compiled for the host with g++ and compared with fp64 autograd
of the harness's PyTorch chain (decoder.RaCFormerSampling / BEVSampling eager paths, which tests/test_decoder.py pins to
the unchanged reference). Tolerance: 2e-4 of the gradient tensor's largest magnitude (fp32 rules vs fp64 autograd)."""
import ctypes
import os
import subprocess

import pytest
import torch

from racformer_b200.decoder import BEVSampling, RaCFormerSampling
from racformer_b200.synthetic import PC_RANGE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_f = ctypes.c_void_p


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ptbwd") / "points_bwd_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "racformer_b200", "csrc"),
                           os.path.join(ROOT, "tests", "csrc", "points_bwd_host.cpp"), "-o", so])
    return ctypes.CDLL(so)


class _CaptureOps:
    """Stands in for SamplingOps: records what the chain hands to the sampling ops."""
    msmv_grouped = None

    def msmv(self, feats, loc, w):
        self.loc, self.w = loc, w
        B_, Q, P, _ = loc.shape
        return loc.new_zeros(B_, Q, 4, P)

    def msda(self, value, shapes, lsi, loc, aw, step):
        self.loc, self.aw = loc, aw
        return loc.new_zeros(loc.shape[0], loc.shape[1], value.shape[2] * value.shape[3])


def _rays(B, Q, g, velocity=True):
    ray = torch.rand(B, Q, 10, generator=g, dtype=torch.float64)
    ray[..., 0] = torch.rand(B, Q, generator=g, dtype=torch.float64)                   # theta in turns
    ray[..., 1] = torch.rand(B, Q, generator=g, dtype=torch.float64) * 0.6 + 0.1       # distance in units of RAY_R
    ray[..., 3:6] = torch.randn(B, Q, 3, generator=g, dtype=torch.float64) * 0.3 + 0.5
    ray[..., 6:8] = torch.randn(B, Q, 2, generator=g, dtype=torch.float64)
    ray[..., 8:] = torch.randn(B, Q, 2, generator=g, dtype=torch.float64) * (0.5 if velocity else 0.0)
    return ray


def _lidar2img(B, T, N, g):
    """Cameras looking outwards in N directions with a plausible intrinsic matrix (704x256 images)."""
    mats = torch.zeros(B, T * N, 4, 4, dtype=torch.float64)
    for b in range(B):
        for i in range(T * N):
            yaw = 2 * torch.pi * (i % N) / N + 0.05 * float(torch.randn((), generator=g))
            c, s = torch.cos(torch.tensor(yaw)), torch.sin(torch.tensor(yaw))
            # lidar -> camera: camera z along (c, s, 0), x to the right, y down
            R = torch.tensor([[float(s), -float(c), 0.0], [0.0, 0.0, -1.0], [float(c), float(s), 0.0]], dtype=torch.float64)
            tvec = torch.randn(3, generator=g, dtype=torch.float64) * 0.3
            K = torch.tensor([[560.0, 0.0, 352.0], [0.0, 560.0, 128.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
            ext = torch.eye(4, dtype=torch.float64)
            ext[:3, :3], ext[:3, 3] = R, tvec
            Kh = torch.eye(4, dtype=torch.float64)
            Kh[:3, :3] = K
            mats[b, i] = Kh @ ext
    return mats


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _close(got, want, name, tol=2e-4):
    scale = float(want.abs().max())
    assert scale > 0, f"{name}: reference gradient is identically zero (test has no power)"
    err = float((got.double() - want).abs().max())
    assert err <= tol * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


def test_image_branch_rules_match_fp64_autograd(host_lib):
    B, Q, T, G, Pn, D, N, L = 2, 7, 3, 2, 2, 3, 3, 4
    g = torch.Generator().manual_seed(5)
    mod = RaCFormerSampling(embed_dims=16, num_frames=T, num_groups=G, num_points=Pn, num_levels=L, depth_num=D, pc_range=PC_RANGE)
    mod.fused_points = False
    P = Pn * D
    ray = _rays(B, Q, g).requires_grad_()
    off = (torch.rand(B, Q, G * P * 3, generator=g, dtype=torch.float64) - 0.5).requires_grad_()
    logit = torch.randn(B, Q, D, generator=g, dtype=torch.float64).requires_grad_()
    sw = torch.randn(B, Q, G * T * P * L, generator=g, dtype=torch.float64)
    time_diff = torch.rand(B, T, generator=g, dtype=torch.float64) * 0.5
    time_diff[:, 0] = 0
    l2i = _lidar2img(B, T, N, g)
    meta = {"time_diff": time_diff, "lidar2img": l2i, "image_h": 256.0, "image_w": 704.0}
    ops = _CaptureOps()
    feats = [torch.zeros(B * T * G, N, 2, 2, 4, dtype=torch.float64)]
    qf = torch.zeros(B, Q, 16, dtype=torch.float64)
    d_region = 0.08
    mod.inner_forward(ops, ray, qf, feats, meta, d_region, heads=(off, logit, sw))
    loc = ops.loc                                                   # [B*T*G,Q,P,3]
    g_loc = torch.randn(loc.shape, generator=g, dtype=torch.float64)
    g_loc[..., 2] = 0
    # like msmv_sampling's backward, hand back no gradient for points nobody sees (their u, v are ill-conditioned)
    seen = ((loc[..., 0] > 0) & (loc[..., 0] < 1) & (loc[..., 1] > 0) & (loc[..., 1] < 1)).unsqueeze(-1)
    g_loc = g_loc * seen
    assert float(seen.double().mean()) > 0.2
    (loc * g_loc).sum().backward()

    f32 = lambda t: t.detach().float().contiguous()
    a = [f32(ray), f32(off), f32(logit), f32(time_diff), f32(l2i),
         torch.linspace(-d_region, d_region, D, dtype=torch.float32)]
    locf, glf = f32(loc), f32(g_loc)
    g_ray, g_off, g_logit = torch.zeros(B, Q, 10), torch.zeros(B, Q, G * P * 3), torch.zeros(B, Q, D)
    pc = (ctypes.c_double * 6)(*PC_RANGE)
    host_lib.host_msmv_points_backward.argtypes = [_f] * 6 + [ctypes.POINTER(ctypes.c_double)] + [ctypes.c_float] * 4 + \
        [ctypes.c_int] * 7 + [_f] * 5
    host_lib.host_msmv_points_backward(*[_ptr(t) for t in a], pc, d_region, 704.0, 256.0, 1e-5, B, Q, T, G, Pn, D, N,
                                       _ptr(locf), _ptr(glf), _ptr(g_ray), _ptr(g_off), _ptr(g_logit))
    _close(g_off, off.grad, "grad_offset")
    _close(g_logit, logit.grad, "grad_ray_logit")
    _close(g_ray, ray.grad, "grad_ray")
    assert float(g_ray[..., 8:].abs().max()) == 0 and float(ray.grad[..., 8:].abs().max()) == 0


def test_bev_branch_rules_match_fp64_autograd(host_lib):
    B, Q, T, M, Pn, D = 2, 9, 3, 4, 2, 3
    g = torch.Generator().manual_seed(6)
    mod = BEVSampling(embed_dims=16, num_frames=T, num_points=Pn, num_heads=M, num_levels=1, pc_range=PC_RANGE, depth_num=D)
    mod.fused_points = False
    P = Pn * D
    ray = _rays(B, Q, g).requires_grad_()
    off = (torch.rand(B, Q, M * P * 2, generator=g, dtype=torch.float64) - 0.5).requires_grad_()
    logit = torch.randn(B, Q, D, generator=g, dtype=torch.float64).requires_grad_()
    sw = torch.randn(B, Q, M * P, generator=g, dtype=torch.float64)
    time_diff = torch.rand(B, T, generator=g, dtype=torch.float64) * 0.5
    time_diff[:, 0] = 0
    meta = {"time_diff": time_diff}
    ops = _CaptureOps()
    value = torch.zeros(B * T, 16, M, 8, dtype=torch.float64)
    qf = torch.zeros(B, Q, M * 8, dtype=torch.float64)
    d_region = 0.08
    mod.sample(ops, ray, qf, value, (4, 4), meta, d_region, heads=(off, logit, sw, torch.zeros(B, Q, T, dtype=torch.float64)), raw=True)
    loc = ops.loc                                                   # [T*B,Q,M,1,P,2]
    g_loc = torch.randn(loc.shape, generator=g, dtype=torch.float64)
    (loc * g_loc).sum().backward()

    f32 = lambda t: t.detach().float().contiguous()
    a = [f32(ray), f32(off), f32(logit), f32(time_diff), torch.linspace(-d_region, d_region, D, dtype=torch.float32)]
    glf = f32(g_loc)
    g_ray, g_off, g_logit = torch.zeros(B, Q, 10), torch.zeros(B, Q, M * P * 2), torch.zeros(B, Q, D)
    pc = (ctypes.c_double * 6)(*PC_RANGE)
    host_lib.host_bev_points_backward.argtypes = [_f] * 5 + [ctypes.POINTER(ctypes.c_double), ctypes.c_float] + \
        [ctypes.c_int] * 6 + [_f] * 4
    host_lib.host_bev_points_backward(*[_ptr(t) for t in a], pc, d_region, B, Q, T, M, Pn, D, _ptr(glf), _ptr(g_ray),
                                      _ptr(g_off), _ptr(g_logit))
    _close(g_off, off.grad, "grad_offset")
    _close(g_logit, logit.grad, "grad_ray_logit")
    _close(g_ray, ray.grad, "grad_ray")
