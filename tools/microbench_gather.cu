// Micro-benchmark behind the design of lsq_scan_kernel: how fast can one B200
// stream 8-byte read records while each record takes (a) one 8-byte gather from a 10 MB table,
// (b) a dependent 16-byte gather from a 5 MB table and (c) one or two 32-bit reductions into a
// 2.4 MB counter table -- all L2 resident, addresses random.  Also measures the FP64 FMA peak that the
// solve roofline is quoted against.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb microbench_gather.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int LEVEL, int ITEMS>
__global__ void __launch_bounds__(256) k_scan(const int32_t *__restrict__ start, const uint8_t *__restrict__ len,
                                             const uint32_t *__restrict__ cnt4, const uint2 *__restrict__ t1, uint32_t n1,
                                             const uint4 *__restrict__ t2, uint32_t n2, uint32_t *ctr, uint32_t n3,
                                             unsigned long long *sink, size_t n) {
    const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long acc = 0;
    for (size_t g = gw; g * (32 * ITEMS) < n; g += nw) {
        const size_t r0 = g * (32 * ITEMS) + (size_t)lane * ITEMS;
        if (r0 + ITEMS > n) break;
        const uint32_t c4 = __ldcs(cnt4 + r0 / 4);
        int32_t s[ITEMS];
        uint32_t l[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            s[k] = __ldcs(start + r0 + k);
            l[k] = __ldcs(len + r0 + k);
        }
        acc += c4;
        if (LEVEL >= 1) {
            uint2 e[ITEMS];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) e[k] = __ldg(t1 + ((uint32_t)s[k] % n1));
            if (LEVEL >= 2) {
                uint4 p[ITEMS];
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) p[k] = __ldg(t2 + ((e[k].x + l[k]) % n2));
                if (LEVEL >= 3) {
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k) {
                        atomicAdd(ctr + (p[k].z % n3), 1u);
                        if (LEVEL >= 4) atomicAdd(ctr + ((p[k].w + 17u) % n3), 1u);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < ITEMS; ++k) acc += p[k].z;
                }
            } else {
#pragma unroll
                for (int k = 0; k < ITEMS; ++k) acc += e[k].x;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) acc += (uint32_t)s[k] + l[k];
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

__global__ void k_dfma(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int LEVEL, int ITEMS>
static void run(const char *name, int ctas_per_sm, const int32_t *start, const uint8_t *len, const uint32_t *cnt4, const uint2 *t1,
                uint32_t n1, const uint4 *t2, uint32_t n2, uint32_t *ctr, uint32_t n3, unsigned long long *sink, size_t n) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = 148 * ctas_per_sm;
    for (int w = 0; w < 2; ++w) k_scan<LEVEL, ITEMS><<<grid, 256>>>(start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int w = 0; w < reps; ++w) k_scan<LEVEL, ITEMS><<<grid, 256>>>(start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("%-44s items %d ctas/sm %d: %.3f ms  %.1f Gread/s  %.2f SM-cycles/read (1.965 GHz)\n", name, ITEMS, ctas_per_sm, ms,
           n / ms * 1e-6, ms * 1e-3 * 148 * 1.965e9 / n);
}

int main(int argc, char **argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 100000000ull;
    const uint32_t n1 = 1250000, n2 = 310000, n3 = 600000;
    std::vector<int32_t> hs(n);
    std::vector<uint8_t> hl(n, 75);
    uint64_t st = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        hs[i] = (int32_t)(st >> 33);
    }
    std::vector<uint2> h1(n1);
    std::vector<uint4> h2(n2);
    for (uint32_t i = 0; i < n1; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h1[i] = make_uint2((uint32_t)(st >> 20), (uint32_t)st); }
    for (uint32_t i = 0; i < n2; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h2[i] = make_uint4((uint32_t)st, 1u, (uint32_t)(st >> 24), (uint32_t)(st >> 11)); }
    int32_t *start; uint8_t *len; uint32_t *cnt4; uint2 *t1; uint4 *t2; uint32_t *ctr; unsigned long long *sink;
    CK(cudaMalloc(&start, n * 4)); CK(cudaMalloc(&len, n)); CK(cudaMalloc(&cnt4, n)); CK(cudaMalloc(&t1, n1 * 8ull));
    CK(cudaMalloc(&t2, n2 * 16ull)); CK(cudaMalloc(&ctr, n3 * 4ull)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemcpy(start, hs.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(len, hl.data(), n, cudaMemcpyHostToDevice));
    CK(cudaMemset(cnt4, 1, n));
    CK(cudaMemcpy(t1, h1.data(), n1 * 8ull, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t2, h2.data(), n2 * 16ull, cudaMemcpyHostToDevice));
    CK(cudaMemset(ctr, 0, n3 * 4ull));
    for (int occ : {4, 8}) {
        run<0, 4>("stream only (6 B/read)", occ, start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
        run<1, 4>("+ 8 B gather (10 MB table)", occ, start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
        run<2, 4>("+ dependent 16 B gather (5 MB table)", occ, start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
        run<3, 4>("+ one red.add.u32 (2.4 MB table)", occ, start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
        run<4, 4>("+ second red.add.u32", occ, start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
    }
    run<3, 8>("two gathers + one red, 8 reads per lane", 4, start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
    run<3, 2>("two gathers + one red, 2 reads per lane", 8, start, len, cnt4, t1, n1, t2, n2, ctr, n3, sink, n);
    // FP64 FMA peak
    double *out;
    CK(cudaMalloc(&out, 148 * 8 * 256 * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 200000;
    k_dfma<<<148 * 8, 256>>>(out, 1000);
    CK(cudaEventRecord(e0));
    k_dfma<<<148 * 8, 256>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("fp64 fma peak: %.2f TFLOP/s (%d threads x %d x 8 FMA)\n", 2.0 * 148 * 8 * 256 * (double)iters * 8 / (ms * 1e-3) / 1e12, 148 * 8 * 256, iters);
    CK(cudaDeviceSynchronize());
    return 0;
}
