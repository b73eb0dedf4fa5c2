// Reverse-mode rules of the fused sampling-point chain (csrc/points.cu forward kernels), per sample point.
//
// The forward chain of one point (models/racformer_transformer.py:361-408 image branch, :493-529 BEV branch;
// models/sparsebev_sampling.py:8-25, 45-120; models/bbox/utils.py:66-106):
//
//   query ray (theta0, d0, z0, log w/l/h, sin, cos, vx, vy) --decode--> box centre (cx,cy,cz), size (w,l,h), yaw (s,c)
//   box-relative offset (ox,oy,oz)  --scale, rotate about z, add centre-->  lidar point (X,Y,Z)
//   velocity warp of frame t (velocity detached), normalise to [0,1]  -->  (xn,yn)
//   cartesian -> polar (theta, dist); dist += linspace[d] + jitter(ray logit)
//   polar -> cartesian, clamp to [0,1]  -->  (x2,y2)               [BEV branch: this is the MSDA sampling location]
//   (x2,y2) -> metric (Xf,Yf), camera projection of the selected view, divide by max(depth, eps) and the image size
//                                                                  [image branch: (u,v) of the MSMV location]
//
// Everything here is plain fp32 arithmetic on values, compiled for the device by nvcc and for the host by g++ (the
// CPU test tests/test_points_bwd_host.py runs exactly these functions against fp64 autograd of the PyTorch chain).
// The discrete view selection is not recomputed: the backward reads the view the forward encoded in loc[...,2].
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define RACF_HD __host__ __device__ __forceinline__
#else
#define RACF_HD inline
#endif

namespace racf {
namespace ptbwd {

constexpr float kMapSize = 102.4f, kCentre = 51.2f, kRayR = 65.0f;
constexpr float kTwoPi = 6.283185307179586f;

struct Consts {
    float pc[6];     // pc_range
    float span[3];   // pc[3+i] - pc[i]
    float d_region;
    int D;           // depth_num
};

struct QueryFrame {   // decode of one query ray, with what its backward needs
    float r0, cos0, sin0;   // polar position
    bool in_x, in_y;        // clamp(.,0,1) pass-through masks of the query centre
    float cx, cy, cz, w, l, h, s, c, vx, vy;
    float sn, cs;           // the raw (sin, cos) regression values
};

struct QueryGrad {    // gradients w.r.t. the decoded frame, summed over the query's points (the chain to the ray is linear)
    float cx, cy, cz, w, l, h, s, c;
};

RACF_HD void zero(QueryGrad& g) { g.cx = g.cy = g.cz = g.w = g.l = g.h = g.s = g.c = 0.f; }

RACF_HD QueryFrame decode(const float* ray, const Consts& k) {
    QueryFrame f;
    const float a0 = ray[0] * kTwoPi;
    f.r0 = ray[1] * kRayR;
    f.cos0 = cosf(a0);
    f.sin0 = sinf(a0);
    const float qx = (kCentre + f.r0 * f.cos0) / kMapSize, qy = (kCentre + f.r0 * f.sin0) / kMapSize;
    f.in_x = qx >= 0.f && qx <= 1.f;
    f.in_y = qy >= 0.f && qy <= 1.f;
    f.cx = fminf(fmaxf(qx, 0.f), 1.f) * k.span[0] + k.pc[0];
    f.cy = fminf(fmaxf(qy, 0.f), 1.f) * k.span[1] + k.pc[1];
    f.cz = ray[2] * k.span[2] + k.pc[2];
    f.w = expf(ray[3]);
    f.l = expf(ray[4]);
    f.h = expf(ray[5]);
    f.sn = ray[6];
    f.cs = ray[7];
    const float ang = atan2f(ray[6], ray[7]);
    f.s = sinf(ang);
    f.c = cosf(ang);
    f.vx = ray[8];
    f.vy = ray[9];
    return f;
}

struct PointFwd {     // forward values of one point that its backward re-uses
    float ex, ey, rho2;          // metric offset from the map centre before the jitter, squared radius
    float r1, cos1, sin1;        // polar position after the jitter
    bool in_x, in_y;             // clamp masks of (x2, y2)
    float x2, y2, Z;             // normalised BEV location, lidar height
    float sig;                   // sigmoid(ray logit)
};

RACF_HD PointFwd point_forward(const QueryFrame& f, float ox, float oy, float oz, float logit, float base, float td,
                               const Consts& k) {
    PointFwd p;
    const float dx = f.w * ox, dy = f.l * oy;
    const float X = f.cx + (dx * f.c - dy * f.s), Y = f.cy + (dx * f.s + dy * f.c);
    p.Z = f.cz + f.h * oz;
    const float xn = (X - f.vx * td - k.pc[0]) / k.span[0], yn = (Y - f.vy * td - k.pc[1]) / k.span[1];
    p.ex = xn * kMapSize - kCentre;
    p.ey = yn * kMapSize - kCentre;
    p.rho2 = p.ex * p.ex + p.ey * p.ey;
    const float dist = sqrtf(p.rho2) / kRayR;
    float t = fmodf(atan2f(p.ey, p.ex) + kTwoPi, kTwoPi);
    if (t < 0.f) t += kTwoPi;
    p.sig = 1.f / (1.f + expf(-logit));
    const float jitter = (p.sig * 2.f - 1.f) * k.d_region / (float)k.D / 2.f;
    p.r1 = (dist + base + jitter) * kRayR;
    p.cos1 = cosf(t);       // cos / sin of theta * 2 pi with theta = t / (2 pi)
    p.sin1 = sinf(t);
    const float x2 = (kCentre + p.r1 * p.cos1) / kMapSize, y2 = (kCentre + p.r1 * p.sin1) / kMapSize;
    p.in_x = x2 >= 0.f && x2 <= 1.f;
    p.in_y = y2 >= 0.f && y2 <= 1.f;
    p.x2 = fminf(fmaxf(x2, 0.f), 1.f);
    p.y2 = fminf(fmaxf(y2, 0.f), 1.f);
    return p;
}

// Camera projection of the image branch (sparsebev_sampling.py:51-75): gradients of (u, v) w.r.t. (x2, y2, Z).
// m: the selected view's lidar2img matrix (row major 4x4).
RACF_HD void project_backward(const PointFwd& p, const float* m, float gu, float gv, float image_w, float image_h, float eps,
                              const Consts& k, float& g_x2, float& g_y2, float& g_Z) {
    const float Xf = p.x2 * k.span[0] + k.pc[0], Yf = p.y2 * k.span[1] + k.pc[1];
    const float camx = m[0] * Xf + m[1] * Yf + m[2] * p.Z + m[3];
    const float camy = m[4] * Xf + m[5] * Yf + m[6] * p.Z + m[7];
    const float camz = m[8] * Xf + m[9] * Yf + m[10] * p.Z + m[11];
    const float hz = fmaxf(camz, eps);
    const float g_camx = gu / (image_w * hz), g_camy = gv / (image_h * hz);
    float g_camz = -(gu * camx / image_w + gv * camy / image_h) / (hz * hz);
    g_camz = camz > eps ? g_camz : (camz == eps ? 0.5f * g_camz : 0.f);     // torch.maximum splits ties
    const float gX = m[0] * g_camx + m[4] * g_camy + m[8] * g_camz;
    const float gY = m[1] * g_camx + m[5] * g_camy + m[9] * g_camz;
    g_Z = m[2] * g_camx + m[6] * g_camy + m[10] * g_camz;
    g_x2 = gX * k.span[0];
    g_y2 = gY * k.span[1];
}

// From the gradient of the normalised location (x2, y2) and of the lidar height Z back to the point's own inputs
// (offset, ray logit) and into the per-query accumulator.
RACF_HD void point_backward(const QueryFrame& f, const PointFwd& p, float ox, float oy, float oz, float g_x2, float g_y2,
                            float g_Z, const Consts& k, float& g_ox, float& g_oy, float& g_oz, float& g_logit, QueryGrad& acc) {
    if (!p.in_x) g_x2 = 0.f;
    if (!p.in_y) g_y2 = 0.f;
    const float g_r1 = (g_x2 * p.cos1 + g_y2 * p.sin1) / kMapSize;
    const float g_t = (g_y2 * p.cos1 - g_x2 * p.sin1) * p.r1 / kMapSize;      // w.r.t. the angle in radians
    const float g_dist = g_r1 * kRayR;
    g_logit = g_dist * (2.f * p.sig * (1.f - p.sig)) * k.d_region / (float)k.D / 2.f;
    float g_ex = 0.f, g_ey = 0.f;
    if (p.rho2 > 0.f) {
        const float rho = sqrtf(p.rho2);
        g_ex = g_dist * p.ex / (rho * kRayR) - g_t * p.ey / p.rho2;
        g_ey = g_dist * p.ey / (rho * kRayR) + g_t * p.ex / p.rho2;
    }
    const float gX = g_ex * kMapSize / k.span[0], gY = g_ey * kMapSize / k.span[1];
    acc.cx += gX;
    acc.cy += gY;
    acc.cz += g_Z;
    const float dx = f.w * ox, dy = f.l * oy;
    const float g_dx = gX * f.c + gY * f.s, g_dy = gY * f.c - gX * f.s;
    acc.c += gX * dx + gY * dy;
    acc.s += gY * dx - gX * dy;
    g_ox = g_dx * f.w;
    g_oy = g_dy * f.l;
    g_oz = g_Z * f.h;
    acc.w += g_dx * ox;
    acc.l += g_dy * oy;
    acc.h += g_Z * oz;
}

// Per-query tail: gradients of the decoded frame -> the ten ray components (velocity is detached in the warp).
RACF_HD void query_backward(const QueryFrame& f, const QueryGrad& g, const Consts& k, float* g_ray) {
    const float g_qx = f.in_x ? g.cx * k.span[0] : 0.f, g_qy = f.in_y ? g.cy * k.span[1] : 0.f;
    const float g_r0 = (g_qx * f.cos0 + g_qy * f.sin0) / kMapSize;
    const float g_a0 = (g_qy * f.cos0 - g_qx * f.sin0) * f.r0 / kMapSize;
    g_ray[0] = g_a0 * kTwoPi;
    g_ray[1] = g_r0 * kRayR;
    g_ray[2] = g.cz * k.span[2];
    g_ray[3] = g.w * f.w;
    g_ray[4] = g.l * f.l;
    g_ray[5] = g.h * f.h;
    const float g_ang = g.s * f.c - g.c * f.s;
    const float n2 = f.sn * f.sn + f.cs * f.cs;
    g_ray[6] = n2 > 0.f ? g_ang * f.cs / n2 : 0.f;
    g_ray[7] = n2 > 0.f ? -g_ang * f.sn / n2 : 0.f;
    g_ray[8] = 0.f;
    g_ray[9] = 0.f;
}

}  // namespace ptbwd
}  // namespace racf
