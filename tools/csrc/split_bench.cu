// Microbenchmark (standalone executable, measurement aid -- not part of the product library): throughput of the exact
// fp32 -> 3 x bf16 operand split on sm_100a, three ways. All variants produce pieces whose sum is the input exactly.
//   0  cvt.rn.bf16x2.f32 (SASS F2FP.BF16.F32.PACK_AB) + shifts + FADD          -- what csrc/mixing_tc.cu does
//   1  Veltkamp splitting on the FMA pipe (x*65537, two subtractions), scalar, PRMT packing
//   2  the same with packed f32x2 arithmetic (mul/sub .f32x2)
//   3  integer round-to-nearest-even (IADD3 / LOP3), PRMT packing
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/csrc/split_bench.cu -o tools/lib/split_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

__device__ __forceinline__ void split_cvt(float a, float b, uint32_t (&q)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        q[k] = *reinterpret_cast<const uint32_t*>(&h);
        if (k < 2) {
            a -= __uint_as_float(q[k] << 16);
            b -= __uint_as_float(q[k] & 0xffff0000u);
        }
    }
}
__device__ __forceinline__ uint32_t pack_hi(float lo, float hi) {      // (hi16 of hi) << 16 | hi16 of lo
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}
__device__ __forceinline__ float veltkamp8(float x) {                  // x rounded to 8 significant bits (a bf16 value)
    const float c = __fmul_rn(x, 65537.0f);
    return __fsub_rn(c, __fsub_rn(c, x));
}
__device__ __forceinline__ void split_velt(float a, float b, uint32_t (&q)[3]) {
    const float a0 = veltkamp8(a), b0 = veltkamp8(b);
    const float ra = __fsub_rn(a, a0), rb = __fsub_rn(b, b0);
    const float a1 = veltkamp8(ra), b1 = veltkamp8(rb);
    q[0] = pack_hi(a0, b0);
    q[1] = pack_hi(a1, b1);
    q[2] = pack_hi(__fsub_rn(ra, a1), __fsub_rn(rb, b1));
}
__device__ __forceinline__ uint64_t pk(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t pack_hi2(uint64_t v) { return __byte_perm((uint32_t)v, (uint32_t)(v >> 32), 0x7632); }
__device__ __forceinline__ void split_velt2(float a, float b, uint32_t (&q)[3]) {
    const uint64_t k = pk(65537.0f, 65537.0f);
    const uint64_t x = pk(a, b);
    uint64_t c = mul2(x, k);
    const uint64_t p0 = sub2(c, sub2(c, x));
    const uint64_t r1 = sub2(x, p0);
    c = mul2(r1, k);
    const uint64_t p1 = sub2(c, sub2(c, r1));
    q[0] = pack_hi2(p0);
    q[1] = pack_hi2(p1);
    q[2] = pack_hi2(sub2(r1, p1));
}
__device__ __forceinline__ float rne8(float x) {
    const uint32_t u = __float_as_uint(x);
    return __uint_as_float((u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u);
}
__device__ __forceinline__ void split_int(float a, float b, uint32_t (&q)[3]) {
    const float a0 = rne8(a), b0 = rne8(b);
    const float ra = a - a0, rb = b - b0;
    const float a1 = rne8(ra), b1 = rne8(rb);
    q[0] = pack_hi(a0, b0);
    q[1] = pack_hi(a1, b1);
    q[2] = pack_hi(ra - a1, rb - b1);
}

template <int kMode>
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ in, uint32_t* __restrict__ out, int iters, long long* cycles) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = in[tid * 8 + j];
    uint32_t acc[3] = {0, 0, 0};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t q[3];
            if (kMode == 0) split_cvt(v[2 * j], v[2 * j + 1], q);
            if (kMode == 1) split_velt(v[2 * j], v[2 * j + 1], q);
            if (kMode == 2) split_velt2(v[2 * j], v[2 * j + 1], q);
            if (kMode == 3) split_int(v[2 * j], v[2 * j + 1], q);
            acc[0] ^= q[0]; acc[1] += q[1]; acc[2] ^= q[2];
            v[2 * j] += 1.0f; v[2 * j + 1] *= 1.0001f;                 // new inputs every round (2 extra FP ops per pair)
        }
    }
    const long long t1 = clock64();
    out[tid * 3] = acc[0]; out[tid * 3 + 1] = acc[1]; out[tid * 3 + 2] = acc[2];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// exactness / equality check of the variants on random data
template <int kMode>
__global__ void check_kernel(const float* __restrict__ in, int n, int* bad_sum, int* diff_vs_cvt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 2 + 1 >= n) return;
    const float a = in[2 * i], b = in[2 * i + 1];
    uint32_t q[3], r[3];
    if (kMode == 1) split_velt(a, b, q);
    if (kMode == 2) split_velt2(a, b, q);
    if (kMode == 3) split_int(a, b, q);
    split_cvt(a, b, r);
    const float sa = (__uint_as_float(q[0] << 16) + __uint_as_float(q[1] << 16)) + __uint_as_float(q[2] << 16);
    const float sb = (__uint_as_float(q[0] & 0xffff0000u) + __uint_as_float(q[1] & 0xffff0000u)) + __uint_as_float(q[2] & 0xffff0000u);
    if (sa != a || sb != b) atomicAdd(bad_sum, 1);
    if (q[0] != r[0] || q[1] != r[1] || q[2] != r[2]) atomicAdd(diff_vs_cvt, 1);
}

int main() {
    const int max_blocks = 148 * 8, threads = 256, iters = 2000;
    const int n = max_blocks * threads * 8;
    std::vector<float> h(n);
    uint32_t s = 12345;
    for (int i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        const float m = (float)((s >> 8) & 0xffffff) / 16777216.0f - 0.5f;
        s = s * 1664525u + 1013904223u;
        h[i] = m * exp2f((float)((int)((s >> 20) & 31) - 16));
    }
    float* in; uint32_t* out; long long* cyc; int* cnt;
    cudaMalloc(&in, n * 4); cudaMalloc(&out, (size_t)max_blocks * threads * 12); cudaMalloc(&cyc, 8); cudaMalloc(&cnt, 8);
    cudaMemcpy(in, h.data(), n * 4, cudaMemcpyHostToDevice);
    printf("{\n");
    // occupancy sweep: warps per SM = 64 (8 CTAs x 8 warps), 8, 4, 1 -- the mixing kernels run the split in 7-16 warps per SM
    const int cfg[4][2] = {{148 * 8, 256}, {148, 256}, {148, 128}, {148, 32}};
    for (int c = 0; c < 4; ++c) {
        const int blocks = cfg[c][0], th = cfg[c][1];
        for (int mode = 0; mode < 4; ++mode) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) split_kernel<0><<<blocks, th>>>(in, out, iters, cyc);
                if (mode == 1) split_kernel<1><<<blocks, th>>>(in, out, iters, cyc);
                if (mode == 2) split_kernel<2><<<blocks, th>>>(in, out, iters, cyc);
                if (mode == 3) split_kernel<3><<<blocks, th>>>(in, out, iters, cyc);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long hcyc = 0;
            cudaMemcpy(&hcyc, cyc, 8, cudaMemcpyDeviceToHost);
            int hc[2] = {0, 0};
            if (mode > 0 && c == 0) {
                cudaMemset(cnt, 0, 8);
                if (mode == 1) check_kernel<1><<<(n / 2 + 255) / 256, 256>>>(in, n, cnt, cnt + 1);
                if (mode == 2) check_kernel<2><<<(n / 2 + 255) / 256, 256>>>(in, n, cnt, cnt + 1);
                if (mode == 3) check_kernel<3><<<(n / 2 + 255) / 256, 256>>>(in, n, cnt, cnt + 1);
                cudaMemcpy(hc, cnt, 8, cudaMemcpyDeviceToHost);
            }
            const double values = (double)blocks * th * 8 * iters;
            printf(" \"warps_per_sm_%d_mode%d\": {\"ms\": %.4f, \"values_per_clk_per_sm\": %.2f, \"cycles_per_8_values_per_warp\": %.1f, "
                   "\"sum_not_exact\": %d, \"pieces_differ_from_cvt\": %d},\n", blocks * th / 32 / 148, mode, ms,
                   values / (ms * 1e-3) / 148.0 / 1.965e9, (double)hcyc / iters, hc[0], hc[1]);
        }
    }
    printf(" \"modes\": \"0 cvt.rn.bf16x2 (F2FP), 1 Veltkamp scalar, 2 Veltkamp f32x2, 3 integer RNE\"\n}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
