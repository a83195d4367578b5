// Micro-benchmark (development tool, not product): how fast can one pass over uint16 EPID frames
// also produce an exact 65536-bin histogram?  Variants:
//   D  stream only (min/max/sum)                         -- upper bound
//   A  global-memory u32 histogram, RED atomics           (many CTAs per frame)
//   B  one CTA per frame, shared-memory packed-u16 histogram (128 KB)
//   C  one CTA per frame, shared-memory u32 histogram of v>>1 (128 KB)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/hist_bench.cu -o tools/hist_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int H0 = 1024, W0 = 1024, CROP = 8, H = H0 - 2 * CROP, W = W0 - 2 * CROP;

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

struct Acc { uint32_t mn, mx; unsigned long long sum; };

__device__ __forceinline__ void acc8(const uint4& q, Acc& a) {
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t lo = w[i] & 0xffff, hi = w[i] >> 16;
        a.mn = min(a.mn, min(lo, hi));
        a.mx = max(a.mx, max(lo, hi));
        a.sum += lo + hi;
    }
}

// ---- D: stream only. grid (tiles, N), each CTA a band of rows
__global__ void k_stream(const uint16_t* __restrict__ frames, int rows_per_cta, unsigned long long* out) {
    const uint16_t* f = frames + (size_t)blockIdx.y * H0 * W0;
    int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, H);
    Acc a{0xffffffffu, 0, 0};
    const int vec_per_row = W / 8;  // 126
    int nvec = (r1 - r0) * vec_per_row;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        int r = r0 + i / vec_per_row, c = i % vec_per_row;
        uint4 q = ldg_stream(reinterpret_cast<const uint4*>(f + (size_t)(r + CROP) * W0 + CROP) + c);
        acc8(q, a);
    }
    if (a.sum == 0x123456789ull) out[0] = a.mn + a.mx;  // keep alive
    atomicAdd(out + 1, a.sum);
}

// ---- A: global hist
__global__ void k_global(const uint16_t* __restrict__ frames, int rows_per_cta, uint32_t* hist, unsigned long long* out) {
    const uint16_t* f = frames + (size_t)blockIdx.y * H0 * W0;
    uint32_t* h = hist + (size_t)blockIdx.y * 65536;
    int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, H);
    Acc a{0xffffffffu, 0, 0};
    const int vec_per_row = W / 8;
    int nvec = (r1 - r0) * vec_per_row;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        int r = r0 + i / vec_per_row, c = i % vec_per_row;
        uint4 q = ldg_stream(reinterpret_cast<const uint4*>(f + (size_t)(r + CROP) * W0 + CROP) + c);
        acc8(q, a);
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { atomicAdd(h + (w[k] & 0xffff), 1u); atomicAdd(h + (w[k] >> 16), 1u); }
    }
    atomicAdd(out + 1, a.sum);
}

// ---- B: CTA per frame, smem packed u16 hist
template <int UNROLL>
__global__ void __launch_bounds__(1024, 1) k_smem16(const uint16_t* __restrict__ frames, int nframes, unsigned long long* out, uint32_t* check) {
    extern __shared__ uint32_t sh[];  // 32768 words
    for (int fidx = blockIdx.x; fidx < nframes; fidx += gridDim.x) {
        for (int i = threadIdx.x; i < 32768; i += blockDim.x) sh[i] = 0;
        __syncthreads();
        const uint16_t* f = frames + (size_t)fidx * H0 * W0;
        Acc a{0xffffffffu, 0, 0};
        const int vec_per_row = W / 8;
        const int nvec = H * vec_per_row;
        for (int i0 = threadIdx.x; i0 < nvec; i0 += blockDim.x * UNROLL) {
            uint4 q[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                int i = i0 + u * blockDim.x;
                if (i < nvec) {
                    int r = i / vec_per_row, c = i % vec_per_row;
                    q[u] = ldg_stream(reinterpret_cast<const uint4*>(f + (size_t)(r + CROP) * W0 + CROP) + c);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                int i = i0 + u * blockDim.x;
                if (i < nvec) {
                    acc8(q[u], a);
                    uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        uint32_t lo = w[k] & 0xffff, hi = w[k] >> 16;
                        atomicAdd(&sh[lo >> 1], (lo & 1) ? 0x10000u : 1u);
                        atomicAdd(&sh[hi >> 1], (hi & 1) ? 0x10000u : 1u);
                    }
                }
            }
        }
        __syncthreads();
        // verification sum of decoded bins
        uint32_t s = 0;
        for (int i = threadIdx.x; i < 32768; i += blockDim.x) s += (sh[i] & 0xffff) + (sh[i] >> 16);
        atomicAdd(check + fidx, s);
        atomicAdd(out + 1, a.sum);
        __syncthreads();
    }
}

// ---- C: CTA per frame, smem u32 hist of v>>1
__global__ void __launch_bounds__(1024, 1) k_smem32(const uint16_t* __restrict__ frames, int nframes, unsigned long long* out, uint32_t* check) {
    extern __shared__ uint32_t sh[];
    for (int fidx = blockIdx.x; fidx < nframes; fidx += gridDim.x) {
        for (int i = threadIdx.x; i < 32768; i += blockDim.x) sh[i] = 0;
        __syncthreads();
        const uint16_t* f = frames + (size_t)fidx * H0 * W0;
        Acc a{0xffffffffu, 0, 0};
        const int vec_per_row = W / 8;
        const int nvec = H * vec_per_row;
        for (int i0 = threadIdx.x; i0 < nvec; i0 += blockDim.x * 4) {
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int i = i0 + u * blockDim.x;
                if (i < nvec) { int r = i / vec_per_row, c = i % vec_per_row; q[u] = ldg_stream(reinterpret_cast<const uint4*>(f + (size_t)(r + CROP) * W0 + CROP) + c); }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int i = i0 + u * blockDim.x;
                if (i < nvec) {
                    acc8(q[u], a);
                    uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                    for (int k = 0; k < 4; k++) { atomicAdd(&sh[(w[k] & 0xffff) >> 1], 1u); atomicAdd(&sh[w[k] >> 17], 1u); }
                }
            }
        }
        __syncthreads();
        uint32_t s = 0;
        for (int i = threadIdx.x; i < 32768; i += blockDim.x) s += sh[i];
        atomicAdd(check + fidx, s);
        atomicAdd(out + 1, a.sum);
        __syncthreads();
    }
}

static float gauss(uint32_t& st) {
    float s = 0;
    for (int i = 0; i < 12; i++) { st = st * 1664525u + 1013904223u; s += (st >> 8) * (1.0f / 16777216.0f); }
    return s - 6.0f;
}

int main(int argc, char** argv) {
    int N = argc > 1 ? atoi(argv[1]) : 296;
    float noise = argc > 2 ? atof(argv[2]) : 130.f;
    size_t fbytes = (size_t)H0 * W0 * 2;
    std::vector<uint16_t> host((size_t)H0 * W0);
    uint32_t st = 12345;
    for (int r = 0; r < H0; r++)
        for (int c = 0; c < W0; c++) {
            float x = (c - 512) / 2.56f;  // mm
            float v = 1500.f;
            float k = roundf(x / 20.f - 0.5f) + 0.5f;
            float d = x - k * 20.f;
            if (fabsf(k) <= 5) v += 52000.f * expf(-0.5f * d * d / (1.6f * 1.6f));
            v += noise * gauss(st);
            host[(size_t)r * W0 + c] = (uint16_t)fminf(fmaxf(v, 0.f), 65535.f);
        }
    uint16_t* d;
    CK(cudaMalloc(&d, fbytes * N));
    for (int i = 0; i < N; i++) CK(cudaMemcpy((char*)d + fbytes * i, host.data(), fbytes, cudaMemcpyHostToDevice));
    uint32_t* hist; CK(cudaMalloc(&hist, (size_t)N * 65536 * 4));
    unsigned long long* out; CK(cudaMalloc(&out, 16));
    uint32_t* check; CK(cudaMalloc(&check, N * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double alg_bytes = (double)N * H * W * 2;
    auto report = [&](const char* name, float ms) { printf("%-28s %8.3f ms  %8.1f GB/s  %9.0f frames/s\n", name, ms, alg_bytes / ms * 1e-6, N / ms * 1e3); };
    int nsm = 148;
    CK(cudaFuncSetAttribute(k_smem16<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(k_smem16<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(k_smem32, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    for (int rep = 0; rep < 3; rep++) {
        float ms;
        for (int rpc : {16, 64}) {
            dim3 g((H + rpc - 1) / rpc, N);
            cudaEventRecord(e0); k_stream<<<g, 256>>>(d, rpc, out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            char nm[64]; snprintf(nm, 64, "D stream rows/cta=%d", rpc); report(nm, ms);
        }
        CK(cudaMemset(hist, 0, (size_t)N * 65536 * 4));
        { dim3 g((H + 15) / 16, N);
          cudaEventRecord(e0); k_global<<<g, 256>>>(d, 16, hist, out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); report("A global RED hist", ms); }
        CK(cudaMemset(check, 0, N * 4));
        cudaEventRecord(e0); k_smem16<4><<<nsm, 1024, 131072>>>(d, N, out, check); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); report("B smem packed-u16 (U4)", ms);
        { uint32_t c0; CK(cudaMemcpy(&c0, check, 4, cudaMemcpyDeviceToHost)); if (c0 != (uint32_t)H * W) printf("   B check mismatch %u vs %u\n", c0, H * W); }
        CK(cudaMemset(check, 0, N * 4));
        cudaEventRecord(e0); k_smem16<8><<<nsm, 1024, 131072>>>(d, N, out, check); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); report("B smem packed-u16 (U8)", ms);
        CK(cudaMemset(check, 0, N * 4));
        cudaEventRecord(e0); k_smem32<<<nsm, 1024, 131072>>>(d, N, out, check); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); report("C smem u32 (v>>1)", ms);
        { uint32_t c0; CK(cudaMemcpy(&c0, check, 4, cudaMemcpyDeviceToHost)); if (c0 != (uint32_t)H * W) printf("   C check mismatch %u vs %u\n", c0, H * W); }
        printf("\n");
    }
    return 0;
}
