// microbench.cu - isolates the two halves of the LUT GEMV on a B200 (standalone, no torch):
//   stream : read a [rows x N] int32 matrix with the kernel's access pattern (strip-major stream-K chunks,
//            STRIPW-column row segments), LDG.128 with U loads in flight per lane, no compute
//   lookup : the 4-bit gather + packed-FMA inner loop on register-resident words (no global traffic),
//            tables and x in shared memory, W warps per CTA, B CTAs per SM
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---------------------------------------------------------------- stream
// Each CTA owns a contiguous chunk of the flattened [strip][row] space; a warp reads (32*4/STRIPW) rows x STRIPW cols per load.
template <int STRIPW, int U>
__global__ void stream_kernel(const uint32_t *__restrict__ q, int rows, int N, int chunk, int T, uint32_t *out) {
    constexpr int LPR = STRIPW / 4;        // lanes per row
    constexpr int RPW = 32 / LPR;          // rows per warp-load
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g0 = min(blockIdx.x * chunk, T), g1 = min(g0 + chunk, T);
    const int li = lane % LPR, lj = lane / LPR;
    uint32_t acc = 0;
    uint4 buf[U];
    auto addr = [&](int g) -> const uint32_t * {
        const int strip = g / rows, r = g - strip * rows;
        return q + (size_t)r * N + strip * STRIPW + 4 * li;
    };
    int g = g0 + warp * RPW + lj;
    const int step = nw * RPW;
#pragma unroll
    for (int u = 0; u < U; ++u) { buf[u] = (g + u * step < g1) ? ldg_stream(addr(g + u * step)) : make_uint4(0, 0, 0, 0); }
    for (; g < g1; g += U * step) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc ^= buf[u].x ^ buf[u].y ^ buf[u].z ^ buf[u].w;
            const int gn = g + (u + U) * step;
            buf[u] = (gn < g1) ? ldg_stream(addr(gn)) : make_uint4(0, 0, 0, 0);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// flat contiguous read (upper bound for any pattern)
template <int U>
__global__ void flat_kernel(const uint4 *__restrict__ q, size_t n4, uint32_t *out) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 buf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) buf[u] = (i + u * stride < n4) ? ldg_stream(q + i + u * stride) : make_uint4(0, 0, 0, 0);
    for (; i < n4; i += U * stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc ^= buf[u].x ^ buf[u].y ^ buf[u].z ^ buf[u].w;
            const size_t in = i + (u + U) * stride;
            buf[u] = (in < n4) ? ldg_stream(q + in) : make_uint4(0, 0, 0, 0);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// ---------------------------------------------------------------- lookup
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_v4(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ffma2(uint64_t &acc, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

template <int MODE>  // 0: PRMT + LDS + FFMA2 (the kernel's loop) ; 1: LDS only (no FMA) ; 2: scalar FFMA instead of FFMA2
__global__ void lookup_kernel(int iters, float *out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)smem_raw + 4095) & ~(uintptr_t)4095);
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(sm);
    float *tab = (float *)sm;
    for (int e = threadIdx.x; e < 1024 + 256; e += blockDim.x) tab[e] = 1.0f + (e & 15) * 0.001f;
    __syncthreads();
    const int lane = threadIdx.x & 31, i16 = lane & 15, jsel = lane >> 4;
    uint32_t ls[4];
    for (int t = 0; t < 4; ++t) ls[t] = (tb & 0xFFFF0000u) | ((((t ^ jsel) << 4) | i16) << 2);
    const uint32_t segc = ((tb >> 8) & 0xF0u) * 0x01010101u;
    const uint32_t xaddr = tb + 4096;
    uint64_t acc[4] = {0, 0, 0, 0};
    float facc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t w0 = threadIdx.x * 2654435761u + blockIdx.x, w1 = w0 * 31 + 7, w2 = w1 * 31 + 7, w3 = w2 * 31 + 7;
    for (int it = 0; it < iters; ++it) {
        const float4 xa = lds_v4(xaddr), xb = lds_v4(xaddr + 16);
        const uint64_t x01 = pack2(xa.x, xa.y), x23 = pack2(xa.z, xa.w), x45 = pack2(xb.x, xb.y), x67 = pack2(xb.z, xb.w);
        uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t E = (w[t] & 0x0F0F0F0Fu) | segc;
            const uint32_t O = ((w[t] >> 4) & 0x0F0F0F0Fu) | segc;
            const float e0 = lds_f32(__byte_perm(E, ls[t], 0x7604)), o0 = lds_f32(__byte_perm(O, ls[t], 0x7604));
            const float e1 = lds_f32(__byte_perm(E, ls[t], 0x7614)), o1 = lds_f32(__byte_perm(O, ls[t], 0x7614));
            const float e2 = lds_f32(__byte_perm(E, ls[t], 0x7624)), o2 = lds_f32(__byte_perm(O, ls[t], 0x7624));
            const float e3 = lds_f32(__byte_perm(E, ls[t], 0x7634)), o3 = lds_f32(__byte_perm(O, ls[t], 0x7634));
            if (MODE == 0) {
                ffma2(acc[t], pack2(e0, o0), x01); ffma2(acc[t], pack2(e1, o1), x23);
                ffma2(acc[t], pack2(e2, o2), x45); ffma2(acc[t], pack2(e3, o3), x67);
            } else if (MODE == 1) {
                facc[t] += e0 + o0 + e1 + o1; facc[t + 4] += e2 + o2 + e3 + o3;
            } else {
                facc[t] = fmaf(e0, xa.x, facc[t]); facc[t + 4] = fmaf(o0, xa.y, facc[t + 4]);
                facc[t] = fmaf(e1, xa.z, facc[t]); facc[t + 4] = fmaf(o1, xa.w, facc[t + 4]);
                facc[t] = fmaf(e2, xb.x, facc[t]); facc[t + 4] = fmaf(o2, xb.y, facc[t + 4]);
                facc[t] = fmaf(e3, xb.z, facc[t]); facc[t + 4] = fmaf(o3, xb.w, facc[t + 4]);
            }
        }
        // new pseudo-random words each iteration (cheap LCG, 4 IMADs per 32 weights)
        w0 = w0 * 1664525u + 1013904223u; w1 = w1 * 1664525u + 1013904223u; w2 = w2 * 1664525u + 1013904223u; w3 = w3 * 1664525u + 1013904223u;
    }
    float s = 0;
    for (int t = 0; t < 4; ++t) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[t])); s += lo + hi; }
    for (int t = 0; t < 8; ++t) s += facc[t];
    if (s == 1234.5f) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int sms = 0, clk = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, max clock %d kHz\n", sms, clk);
    uint32_t *dout; CK(cudaMalloc(&dout, 64));
    // big buffer: 1 GiB, so every pass streams from HBM
    const size_t bytes = 1ull << 30;
    uint32_t *q; CK(cudaMalloc(&q, bytes)); CK(cudaMemset(q, 1, bytes));

    {   // does reserving shared memory (smaller L1) or extra idle warps slow the LDG stream?
        const int N = 4096;
        const int rows = (int)(bytes / 4 / N) / 16 * 16;
        const double mb = (double)rows * N * 4;
        const int strips = N / 64; const long long T = (long long)strips * rows;
        CK(cudaFuncSetAttribute(stream_kernel<64, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        for (int occ : {2, 3}) for (int smem_kb : {0, 16, 32, 48, 64, 72}) for (int threads : {256, 288}) {
            const int G = sms * occ; const int chunk = (int)((T + G - 1) / G);
            float ms = time_ms([&] { stream_kernel<64, 6><<<G, threads, smem_kb * 1024>>>(q, rows, N, chunk, (int)T, dout); }, 3);
            printf("stream U=6 CTAs/SM=%d threads=%d smem/CTA=%3d KB : %.1f GB/s\n", occ, threads, smem_kb, mb / ms / 1e6);
        }
    }
    return 0;
}
