// Host build of the per-point backward rules of racformer_b200/csrc/points_bwd.cuh (test infrastructure): the same
// functions the CUDA kernels of csrc/points_train.cu call, driven by serial loops, so that the CPU suite can check the
// hand-derived gradients against fp64 autograd of the PyTorch chain (tests/test_points_bwd_host.py).
#include "points_bwd.cuh"

using namespace racf::ptbwd;

static Consts make_consts(const double* pc_range, float d_region, int D) {
    Consts k;
    for (int i = 0; i < 6; ++i) k.pc[i] = (float)pc_range[i];
    for (int i = 0; i < 3; ++i) k.span[i] = (float)(pc_range[3 + i] - pc_range[i]);
    k.d_region = d_region;
    k.D = D;
    return k;
}

extern "C" void host_msmv_points_backward(const float* ray, const float* offset, const float* ray_logit, const float* time_diff,
                                          const float* lidar2img, const float* depth_base, const double* pc_range,
                                          float d_region, float image_w, float image_h, float eps, int B, int Q, int T, int G,
                                          int Pn, int D, int N, const float* loc, const float* grad_loc, float* g_ray,
                                          float* g_offset, float* g_logit) {
    const Consts k = make_consts(pc_range, d_region, D);
    const int P = Pn * D, GP = G * P;
    for (long long bq = 0; bq < (long long)B * Q; ++bq) {
        const int b = (int)(bq / Q), q = (int)(bq % Q);
        const QueryFrame f = decode(ray + bq * 10, k);
        QueryGrad acc;
        zero(acc);
        for (int d = 0; d < D; ++d) g_logit[bq * D + d] = 0.f;
        for (int gp = 0; gp < GP; ++gp) {
            const int g = gp / P, p = gp % P, d = p % D;
            const float* off = offset + (bq * GP + gp) * 3;
            float so[3] = {0.f, 0.f, 0.f};
            for (int t = 0; t < T; ++t) {
                const PointFwd pf = point_forward(f, off[0], off[1], off[2], ray_logit[bq * D + d], depth_base[d],
                                                  time_diff[b * T + t], k);
                const long long row = ((((long long)b * T + t) * G + g) * Q + q) * P + p;
                const int view = (int)lrintf(loc[row * 3 + 2] * (float)(N - 1));
                const float* m = lidar2img + (((long long)b * T + t) * N + view) * 16;
                float g_x2, g_y2, g_Z, go[3], gl;
                project_backward(pf, m, grad_loc[row * 3], grad_loc[row * 3 + 1], image_w, image_h, eps, k, g_x2, g_y2, g_Z);
                point_backward(f, pf, off[0], off[1], off[2], g_x2, g_y2, g_Z, k, go[0], go[1], go[2], gl, acc);
                for (int i = 0; i < 3; ++i) so[i] += go[i];
                g_logit[bq * D + d] += gl;
            }
            for (int i = 0; i < 3; ++i) g_offset[(bq * GP + gp) * 3 + i] = so[i];
        }
        query_backward(f, acc, k, g_ray + bq * 10);
    }
}

extern "C" void host_bev_points_backward(const float* ray, const float* offset, const float* ray_logit, const float* time_diff,
                                         const float* depth_base, const double* pc_range, float d_region, int B, int Q, int T,
                                         int M, int Pn, int D, const float* grad_loc, float* g_ray, float* g_offset,
                                         float* g_logit) {
    const Consts k = make_consts(pc_range, d_region, D);
    const int P = Pn * D, MP = M * P;
    const long long frame_stride = (long long)B * Q * MP;
    for (long long bq = 0; bq < (long long)B * Q; ++bq) {
        const int b = (int)(bq / Q);
        const QueryFrame f = decode(ray + bq * 10, k);
        QueryGrad acc;
        zero(acc);
        for (int d = 0; d < D; ++d) g_logit[bq * D + d] = 0.f;
        for (int mp = 0; mp < MP; ++mp) {
            const int d = (mp % P) % D;
            const float* off = offset + (bq * MP + mp) * 2;
            float so[2] = {0.f, 0.f};
            for (int t = 0; t < T; ++t) {
                const PointFwd pf = point_forward(f, off[0], off[1], 0.f, ray_logit[bq * D + d], depth_base[d],
                                                  time_diff[b * T + t], k);
                const long long o = (long long)t * frame_stride + bq * MP + mp;
                float gx, gy, gz, gl;
                point_backward(f, pf, off[0], off[1], 0.f, grad_loc[o * 2], grad_loc[o * 2 + 1], 0.f, k, gx, gy, gz, gl, acc);
                so[0] += gx;
                so[1] += gy;
                g_logit[bq * D + d] += gl;
            }
            g_offset[(bq * MP + mp) * 2] = so[0];
            g_offset[(bq * MP + mp) * 2 + 1] = so[1];
        }
        query_backward(f, acc, k, g_ray + bq * 10);
    }
}
