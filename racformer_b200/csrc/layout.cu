// Channel-last re-layout of one FPN level for the sampling op -- SURVEY.md 8f-3.
//
// The reference decoder does `feat.reshape(B,T,N,G,C,H,W).permute(0,1,3,2,5,6,4).reshape(B*T*G,N,H,W,C).contiguous()`
// (models/racformer_transformer.py:112-124) through PyTorch's generic strided copy (measured 0.87 ms for the 735 MB
// f8 pyramid on B200, 1.7 TB/s). This is the same copy as a tiled transpose: each (b,t,n,g) slab is a [C][H*W] matrix
// that becomes [H*W][C]; 64 channels x 32 pixels go through shared memory so that both the reads (along pixels) and
// the writes (along channels, 256 B per pixel) are coalesced.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <stdint.h>

#include "racf_common.cuh"
#include "racformer_ops.h"

namespace racf {

constexpr int kTilePix = 32;

__device__ __forceinline__ float load_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_as_float(const __half* p) { return __half2float(__ldg(p)); }

template <typename In>
__global__ void __launch_bounds__(256) to_sampling_layout_c64_kernel(const In* __restrict__ in, float* __restrict__ out,
                                                                     int T, int N, int G, int HW) {
    __shared__ float tile[64][kTilePix + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (HW + kTilePix - 1) / kTilePix;
    const long long slab = blockIdx.x / tiles;            // ((b*T + t)*N + n)*G + g  in INPUT order
    const int p0 = (blockIdx.x % tiles) * kTilePix;
    const int g = (int)(slab % G);
    const long long btn = slab / G;
    const int n = (int)(btn % N);
    const long long bt = btn / N;
    const In* src = in + (btn * G + g) * 64 * (long long)HW;                    // [C][HW]
    float* dst = out + ((bt * G + g) * N + n) * (long long)HW * 64;             // [HW][C]
    const int p = p0 + lane;
#pragma unroll
    for (int c = warp; c < 64; c += 8) tile[c][lane] = (p < HW) ? load_as_float(src + (long long)c * HW + p) : 0.f;
    __syncthreads();
#pragma unroll
    for (int q = warp; q < kTilePix; q += 8) {
        if (p0 + q < HW) {
            float* row = dst + (long long)(p0 + q) * 64;
            row[lane] = tile[lane][q];
            row[lane + 32] = tile[lane + 32][q];
        }
    }
}

// The inverse copy (backward of the re-layout): [HW][64] slabs of the sampling layout -> [64][HW] slabs of the FPN layout.
__global__ void __launch_bounds__(256) from_sampling_layout_c64_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                       int T, int N, int G, int HW) {
    __shared__ float tile[kTilePix][64 + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (HW + kTilePix - 1) / kTilePix;
    const long long slab = blockIdx.x / tiles;            // ((b*T + t)*N + n)*G + g  in FPN order
    const int p0 = (blockIdx.x % tiles) * kTilePix;
    const int g = (int)(slab % G);
    const long long btn = slab / G;
    const int n = (int)(btn % N);
    const long long bt = btn / N;
    const float* src = in + ((bt * G + g) * N + n) * (long long)HW * 64;        // [HW][C]
    float* dst = out + (btn * G + g) * 64 * (long long)HW;                       // [C][HW]
#pragma unroll
    for (int q = warp; q < kTilePix; q += 8) {
        const bool ok = p0 + q < HW;
        const float* row = src + (long long)(p0 + q) * 64;
        tile[q][lane] = ok ? __ldg(row + lane) : 0.f;
        tile[q][lane + 32] = ok ? __ldg(row + lane + 32) : 0.f;
    }
    __syncthreads();
    const int p = p0 + lane;
#pragma unroll
    for (int c = warp; c < 64; c += 8)
        if (p < HW) dst[(long long)c * HW + p] = tile[lane][c];
}

// fp16 channel-last input (the FPN's own output format under mixed precision: NHWC halves, models/racformer.py:106
// `auto_fp16(..., out_fp32=True)` casts exactly these values up) -> the fp32 sampling layout, upcast fused into the copy.
// in [B*T*N][HW][G*64] halves; one thread moves 8 channels: a 16-byte load, two 16-byte stores.
__global__ void __launch_bounds__(256) nhwc_half_to_sampling_layout_c64_kernel(const uint4* __restrict__ in, float4* __restrict__ out,
                                                                               int N, int G, int HW, long long total8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i & 7);
        long long r = i >> 3;
        const int g = (int)(r % G);
        r /= G;
        const int p = (int)(r % HW);
        const long long btn = r / HW;
        const int n = (int)(btn % N);
        const long long bt = btn / N;
        const uint4 v = __ldg(in + i);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
        float4* dst = out + ((((bt * G + g) * N + n) * (long long)HW + p) * 16 + c8 * 2);
        dst[0] = make_float4(a.x, a.y, b.x, b.y);
        dst[1] = make_float4(c.x, c.y, d.x, d.y);
    }
}

// in [batch][C][S] -> out[batch][S][ld] (first C of ld floats per pixel) and, optionally, the same into out2 with its
// own row length: the NCHW -> NHWC conversion in front of the radar temporal encoder's convolutions, written once into
// the dense channel-last tensor and once into the first C channels of the [C + hidden]-channel concatenation buffer
// (PyTorch's generic strided copy takes 153 us for the 134 MB f8 BEV queue, this tiled transpose about a third).
__global__ void __launch_bounds__(256) chw_to_hwc_kernel(const float* __restrict__ in, float* __restrict__ out, int ld,
                                                         float* __restrict__ out2, int ld2, int C, int S) {
    __shared__ float tile[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const long long b = blockIdx.z;
    const float* src = in + b * (long long)C * S;
#pragma unroll
    for (int c = warp; c < 32; c += 8)
        tile[c][lane] = (c0 + c < C && s0 + lane < S) ? __ldg(src + (long long)(c0 + c) * S + s0 + lane) : 0.f;
    __syncthreads();
#pragma unroll
    for (int q = warp; q < 32; q += 8) {
        if (s0 + q < S && c0 + lane < C) {
            const float v = tile[lane][q];
            out[(b * S + s0 + q) * ld + c0 + lane] = v;
            if (out2 != nullptr) out2[(b * S + s0 + q) * ld2 + c0 + lane] = v;
        }
    }
}

// ConvGRU cell gates of the radar temporal encoder (models/racformer_transformer.py:709-720) on channel-last tensors:
//   gates [P pixels][3 Hc] = (z | r | cand) pre-activations (the gates convolution's output), h_prev [P][Hc] ->
//   h = (1 - sigmoid(z)) h_prev + sigmoid(z) tanh(cand + sigmoid(r) h_prev)                          [P][Hc]
// One launch instead of nine PyTorch elementwise kernels per recurrence step; separate multiplies / adds where PyTorch
// runs separate kernels (no FMA contraction), the same libm calls.
__global__ void __launch_bounds__(256) convgru_gates_kernel(const float* __restrict__ gates, const float* __restrict__ h_prev,
                                                            float* __restrict__ h, long long quads, int hc4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const long long pix = i / hc4;
    const int c4 = (int)(i - pix * hc4);
    const float4* g = reinterpret_cast<const float4*>(gates) + pix * 3 * hc4 + c4;
    const float4 zv = __ldg(g), rv = __ldg(g + hc4), cv = __ldg(g + 2 * hc4);
    const float4 hv = __ldg(reinterpret_cast<const float4*>(h_prev) + i);
    auto cell = [](float zl, float rl, float cl, float hp) {
        const float z = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-zl)));
        const float r = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-rl)));
        const float cand = tanhf(__fadd_rn(cl, __fmul_rn(r, hp)));
        return __fadd_rn(__fmul_rn(__fsub_rn(1.f, z), hp), __fmul_rn(z, cand));
    };
    reinterpret_cast<float4*>(h)[i] = make_float4(cell(zv.x, rv.x, cv.x, hv.x), cell(zv.y, rv.y, cv.y, hv.y),
                                                  cell(zv.z, rv.z, cv.z, hv.z), cell(zv.w, rv.w, cv.w, hv.w));
}

// nn.Upsample(scale_factor = 2, mode = "bilinear", align_corners = True) on a channel-last tensor (the radar temporal encoder's
// hidden state, models/racformer_transformer.py:637-640): in [N, H, W, C] -> out [N, 2H, 2W, C]. A thread owns 4 channels of one
// output pixel; source coordinate = dst * (in - 1) / (out - 1), the same interpolation weights as ATen's kernel.
__global__ void __launch_bounds__(256) upsample2x_bilinear_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int H,
                                                                       int W, int c4, long long quads) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const int HO = 2 * H, WO = 2 * W;
    long long r = i;
    const int c = (int)(r % c4); r /= c4;
    const int x2 = (int)(r % WO); r /= WO;
    const int y2 = (int)(r % HO);
    const long long n = r / HO;
    const float rh = HO > 1 ? (float)(H - 1) / (float)(HO - 1) : 0.f, rw = WO > 1 ? (float)(W - 1) / (float)(WO - 1) : 0.f;
    const float h1r = rh * (float)y2, w1r = rw * (float)x2;
    const int h1 = (int)h1r, w1 = (int)w1r;
    const int hp = h1 < H - 1 ? 1 : 0, wp = w1 < W - 1 ? 1 : 0;
    const float hl1 = h1r - (float)h1, hl0 = 1.f - hl1, wl1 = w1r - (float)w1, wl0 = 1.f - wl1;
    const float4* src = reinterpret_cast<const float4*>(in) + ((n * H + h1) * W + w1) * c4 + c;
    const float4 v00 = __ldg(src), v01 = __ldg(src + (long long)wp * c4), v10 = __ldg(src + (long long)hp * W * c4),
                 v11 = __ldg(src + ((long long)hp * W + wp) * c4);
    auto mix = [&](float a, float b, float cc, float d) { return hl0 * (wl0 * a + wl1 * b) + hl1 * (wl0 * cc + wl1 * d); };
    reinterpret_cast<float4*>(out)[i] = make_float4(mix(v00.x, v01.x, v10.x, v11.x), mix(v00.y, v01.y, v10.y, v11.y),
                                                    mix(v00.z, v01.z, v10.z, v11.z), mix(v00.w, v01.w, v10.w, v11.w));
}

}  // namespace racf

extern "C" int racf_upsample2x_bilinear_nhwc(const float* in, int batch, int height, int width, int channels, float* out,
                                             racf_stream_t stream) {
    if (!in || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || height <= 0 || width <= 0 || channels <= 0) return RACF_ERR_BAD_SHAPE;
    if ((channels & 3) != 0 || ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u)) return RACF_ERR_UNSUPPORTED;
    const long long quads = (long long)batch * 4 * height * width * (channels / 4);
    const long long blocks = (quads + 255) / 256;
    if (blocks >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    racf::upsample2x_bilinear_nhwc_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, height, width,
                                                                                                          channels / 4, quads);
    return (int)cudaGetLastError();
}

extern "C" int racf_convgru_gates_forward(const float* gates, const float* h_prev, long long pixels, int hidden_channels,
                                          float* h, racf_stream_t stream) {
    if (!gates || !h_prev || !h) return RACF_ERR_NULL_POINTER;
    if (pixels <= 0 || hidden_channels <= 0) return RACF_ERR_BAD_SHAPE;
    if ((hidden_channels & 3) != 0 || ((reinterpret_cast<uintptr_t>(gates) | reinterpret_cast<uintptr_t>(h_prev) |
                                        reinterpret_cast<uintptr_t>(h)) & 15u))
        return RACF_ERR_UNSUPPORTED;
    const long long quads = pixels * (hidden_channels / 4);
    const long long blocks = (quads + 255) / 256;
    if (blocks >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    racf::convgru_gates_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(gates, h_prev, h, quads,
                                                                                               hidden_channels / 4);
    return (int)cudaGetLastError();
}

extern "C" int racf_chw_to_hwc(const float* in, int batch, int channels, int positions, float* out, int ld, float* out2,
                               int ld2, racf_stream_t stream) {
    if (!in || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || channels <= 0 || positions <= 0 || batch > 65535 || ld < channels || (out2 && ld2 < channels))
        return RACF_ERR_BAD_SHAPE;
    const dim3 grid((unsigned)((positions + 31) / 32), (unsigned)((channels + 31) / 32), (unsigned)batch);
    if (grid.y > 65535) return RACF_ERR_BAD_SHAPE;
    racf::chw_to_hwc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, ld, out2, ld2, channels, positions);
    return (int)cudaGetLastError();
}

// in: [B, T*N, G*64, H, W] contiguous -> out: [B*T*G, N, H, W, 64] contiguous. Only C == 64 is implemented.
extern "C" int racf_to_sampling_layout(const float* in, float* out, int batch, int num_frames, int num_views,
                                       int num_groups, int channels, int height, int width, racf_stream_t stream) {
    if (!in || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_frames <= 0 || num_views <= 0 || num_groups <= 0 || height <= 0 || width <= 0 || channels != 64)
        return RACF_ERR_BAD_SHAPE;
    const int HW = height * width;
    const long long slabs = (long long)batch * num_frames * num_views * num_groups;
    const long long blocks = slabs * ((HW + racf::kTilePix - 1) / racf::kTilePix);
    if (blocks >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    racf::to_sampling_layout_c64_kernel<float><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, num_frames, num_views, num_groups, HW);
    return (int)cudaGetLastError();
}

extern "C" int racf_from_sampling_layout(const float* in, float* out, int batch, int num_frames, int num_views,
                                         int num_groups, int channels, int height, int width, racf_stream_t stream) {
    if (!in || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_frames <= 0 || num_views <= 0 || num_groups <= 0 || height <= 0 || width <= 0 || channels != 64)
        return RACF_ERR_BAD_SHAPE;
    const int HW = height * width;
    const long long slabs = (long long)batch * num_frames * num_views * num_groups;
    const long long blocks = slabs * ((HW + racf::kTilePix - 1) / racf::kTilePix);
    if (blocks >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    racf::from_sampling_layout_c64_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, num_frames, num_views, num_groups, HW);
    return (int)cudaGetLastError();
}

// fp16 input, fp32 sampling layout out. channels_last == 0: in [B, T*N, G*64, H, W] halves (the tiled transpose above with
// 2-byte loads); channels_last != 0: in [B*T*N, H, W, G*64] halves (16-byte aligned).
extern "C" int racf_to_sampling_layout_f16(const void* in, float* out, int batch, int num_frames, int num_views,
                                           int num_groups, int channels, int height, int width, int channels_last,
                                           racf_stream_t stream) {
    if (!in || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_frames <= 0 || num_views <= 0 || num_groups <= 0 || height <= 0 || width <= 0 || channels != 64)
        return RACF_ERR_BAD_SHAPE;
    const int HW = height * width;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (channels_last) {
        if ((reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return RACF_ERR_UNSUPPORTED;
        const long long total8 = (long long)batch * num_frames * num_views * HW * num_groups * 8;
        const long long cap = 32LL * racf::sm_count(), want = (total8 + 255) / 256;
        racf::nhwc_half_to_sampling_layout_c64_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(
            static_cast<const uint4*>(in), reinterpret_cast<float4*>(out), num_views, num_groups, HW, total8);
        return (int)cudaGetLastError();
    }
    const long long slabs = (long long)batch * num_frames * num_views * num_groups;
    const long long blocks = slabs * ((HW + racf::kTilePix - 1) / racf::kTilePix);
    if (blocks >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    racf::to_sampling_layout_c64_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const __half*>(in), out, num_frames,
                                                                                  num_views, num_groups, HW);
    return (int)cudaGetLastError();
}
