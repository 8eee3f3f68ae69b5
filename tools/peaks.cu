// peaks.cu -- micro-benchmarks that measure the roofline denominators MEASURED_PEAKS.json does not hold:
//   * FP64 pipe issue rate with unfused DADD/DMUL (the bit-exact pair kernel cannot use FMA) and with DFMA
//   * shared-memory atomic-add throughput for spread addresses (the histogram update)
//   * streaming read bandwidth with 16-byte loads (the MSD kernel's access pattern)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu
// Run (on the GPU box): tools/peaks > gpurun_out/peaks.json
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess) {                                                                 \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));           \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

template <int MODE>   // 0: DADD+DMUL unfused alternating, 1: DFMA (two uniform operands), 2: DFMA with three register operands
__global__ void __launch_bounds__(256) k_fp64(double *out, double a, double b)
{
    double v[CHAINS], w[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
        v[k] = a + threadIdx.x * 1e-9 + k;
        w[k] = b * (threadIdx.x + k + 1);
    }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (MODE == 2) {
                // the correlation kernels' shape: accumulator += window[k'] * broadcast value, all three in registers
                v[k] = fma(w[(k + 1) % CHAINS], w[0], v[k]);
                v[k] = fma(w[(k + 3) % CHAINS], w[1], v[k]);
            } else if (MODE == 0) {
                v[k] = __dadd_rn(v[k], b);
                v[k] = __dmul_rn(v[k], a);
            } else {
                v[k] = fma(v[k], a, b);
                v[k] = fma(v[k], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += v[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_smem_atomics(unsigned *out, int nb, int iters)
{
    extern __shared__ unsigned h[];
    for (int k = threadIdx.x; k < nb; k += blockDim.x) h[k] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(&h[(x >> 8) % nb], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0 && h[0] == 0xffffffffu) out[0] = h[1];
}

__global__ void __launch_bounds__(256) k_stream_read(const double2 *p, size_t n2, double *out)
{
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 v;
        asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + i));
        s += v.x + v.y;
    }
    if (s == 12345.678) out[0] = s;
}

// ---- accuracy of MUFU.SQRT (sqrt.approx.ftz.f32), every positive normal fp32 input: k_pair_fast's error bound
// (csrc/pair_fast.cuh) budgets 2^-22 for it.  out[0] = max over all inputs of |approx - sqrt| / sqrt as ulps-of-2^-24, scaled.
__global__ void __launch_bounds__(256) k_mufu_sqrt_err(unsigned long long *worst)
{
    // inputs: bit patterns 0x00800000 .. 0x7f7fffff (positive normals), grid-strided
    double w = 0.0;
    for (unsigned long long b = 0x00800000ull + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= 0x7f7fffffull;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)b);
        float a;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(x));
        const double exact = sqrt((double)x);
        const double rel = fabs((double)a - exact) / exact;
        w = rel > w ? rel : w;
    }
    // max-reduce through the order-preserving bit pattern of a non-negative double
    atomicMax(worst, (unsigned long long)__double_as_longlong(w));
}

// FP32 issue rate (FFMA), and the integer POPC rate the popcount survival kernel is bounded by
__global__ void __launch_bounds__(256) k_fp32(float *out, float a, float b)
{
    float v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = a + (float)(threadIdx.x + c);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) v[c] = __fmaf_rn(v[c], a, b);
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) v[c] = __fmaf_rn(v[c], b, a);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += v[c];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_popc(unsigned long long *out, unsigned long long a)
{
    unsigned long long v[CHAINS];
    unsigned int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) v[c] = a * (threadIdx.x + c + 1);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            acc += __popcll(v[c] & (v[c] >> (it & 63)));
            v[c] += acc;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <class F>
float time_ms(F f, int reps)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *d_out;
    CK(cudaMalloc(&d_out, 1024));
    const int blocks = sms * 8;
    const double ops = (double)blocks * 256 * ITERS * CHAINS * 2;
    float t0 = time_ms([&] { k_fp64<0><<<blocks, 256>>>(d_out, 1.0000001, 1e-7); }, 5);
    float t1 = time_ms([&] { k_fp64<1><<<blocks, 256>>>(d_out, 1.0000001, 1e-7); }, 5);
    float t2 = time_ms([&] { k_fp64<2><<<blocks, 256>>>(d_out, 1.0000001, 1e-7); }, 5);
    // sustained: repeat for ~2 s to see the clock settle under the power cap
    float t0s = 0;
    {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        const int reps = (int)(2000.0f / t0) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) k_fp64<0><<<blocks, 256>>>(d_out, 1.0000001, 1e-7);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&t0s, e0, e1));
        t0s /= reps;
    }
    const int nb = 400, aiters = 4096;
    CK(cudaFuncSetAttribute(k_smem_atomics, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    float ta = time_ms([&] { k_smem_atomics<<<sms * 4, 256, nb * 4>>>((unsigned *)d_out, nb, aiters); }, 5);
    float ta2 = time_ms([&] { k_smem_atomics<<<sms * 4, 256, 6000 * 4>>>((unsigned *)d_out, 6000, aiters); }, 5);
    const size_t bytes = (size_t)4 << 30;
    double2 *buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    float tr = time_ms([&] { k_stream_read<<<sms * 16, 256>>>(buf, bytes / 16, d_out); }, 5);
    // MUFU.SQRT accuracy over every positive normal fp32 input
    unsigned long long *d_worst;
    CK(cudaMalloc(&d_worst, 8));
    CK(cudaMemset(d_worst, 0, 8));
    k_mufu_sqrt_err<<<sms * 16, 256>>>(d_worst);
    CK(cudaDeviceSynchronize());
    unsigned long long hw = 0;
    CK(cudaMemcpy(&hw, d_worst, 8, cudaMemcpyDeviceToHost));
    double worst_rel;
    memcpy(&worst_rel, &hw, 8);
    float tf32 = time_ms([&] { k_fp32<<<blocks, 256>>>((float *)d_out, 1.0000001f, 1e-7f); }, 5);
    float tpop = time_ms([&] { k_popc<<<blocks, 256>>>((unsigned long long *)d_out, 0x9e3779b97f4a7c15ull); }, 5);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d,\n", prop.name, sms, prop.clockRate / 1000);
    printf(" \"mufu_sqrt_max_rel_err\": %.6e, \"mufu_sqrt_max_rel_err_log2\": %.3f,\n", worst_rel, log2(worst_rel));
    printf(" \"fp32_ffma_tflops_burst\": %.3f, \"popc64_and_shift_gops\": %.1f,\n", 2 * ops / tf32 * 1e-9,
           (double)blocks * 256 * ITERS * CHAINS / tpop * 1e-6);
    printf(" \"fp64_unfused_tflops_burst\": %.3f, \"fp64_unfused_tflops_sustained\": %.3f, \"fp64_fma_tflops_burst\": %.3f,\n",
           ops / t0 * 1e-9, ops / t0s * 1e-9, 2 * ops / t1 * 1e-9);
    printf(" \"fp64_fma_3reg_tflops_burst\": %.3f,\n", 2 * ops / t2 * 1e-9);
    printf(" \"smem_atomic_gops_400bins\": %.2f, \"smem_atomic_gops_6000bins\": %.2f,\n",
           (double)sms * 4 * 256 * aiters / ta * 1e-6, (double)sms * 4 * 256 * aiters / ta2 * 1e-6);
    printf(" \"stream_read_gbs\": %.1f,\n", bytes / tr * 1e-6);
    printf(" \"how\": \"8 independent chains/thread of DADD+DMUL (unfused) or DFMA, %d blocks x 256 thr, best of 5; sustained = back-to-back for 2 s; smem atomics = random bins, 4 CTAs/SM; read = 4 GiB ld.global.cs.v2.f64\"}\n",
           blocks);
    return 0;
}
