// Microbenchmark: the x-transform of the Poisson solve (64 rows x 96-point real DFT, forward or inverse) as a tensor-core GEMM
//   out[64][96] = in[64][96] * F[96][96]          (F = the real DFT matrix: 49 cosine and 47 sine columns)
// on sm_100a: tcgen05.mma kind::tf32 (M = 64, N = 96, K = 8 per instruction, 12 instructions per product), operands in shared
// memory (K-major, no swizzle: 8 x 16 B core matrices), fp32 accumulator in tensor memory, read back with tcgen05.ld.
// TF32 keeps 10 mantissa bits, the fp32 kernel needs ~24: the standard remedy is the 3xTF32 split  a = a_hi + a_lo,
// out ~= a_hi F_hi + a_lo F_hi + a_hi F_lo  (three products accumulated in tensor memory, 36 instructions).
//
// What is timed, per CTA (one CTA per SM, 384 threads like the product kernel, clock64 of thread 0, averaged over `iters`):
//   split   : 384 threads turn a row-major fp32 tile in shared memory into the hi/lo tf32 core-matrix tiles
//   mma1/3  : issue of 12 / 36 tcgen05.mma by one thread + tcgen05.commit -> mbarrier wait (the tensor-core time)
//   ldtm    : 4 warps read the 64 x 96 accumulator from tensor memory and store it row-major to shared memory
// and the error of both variants against an fp64 DFT.  The numbers go into DESIGN.md section 4.1 next to the cycles of the
// in-register FFT passes they would replace (tools/ncu_summary.py of the product kernel).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench/dft_gemm tools/microbench/dft_gemm.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int M = 64, N = 96, K = 96;
constexpr int NT = 384;
constexpr int A_RG = M / 8, B_RG = N / 8, KC = K / 4;            // row groups of 8, K chunks of 4 tf32 (16 bytes)
constexpr int A_BYTES = M * K * 4, B_BYTES = N * K * 4;
constexpr int TMEM_COLS = 128;                                   // power of two >= N
constexpr int RS = 100;                                          // padded row of the row-major tile: 16-byte accesses of 8 consecutive rows hit 8 different bank groups

// element (row r, column k) of a K-major, unswizzled operand: core matrix (k/4, r/8) of 128 contiguous bytes
__host__ __device__ inline int core_off(int r, int k, int nrg) { return ((k >> 2) * nrg + (r >> 3)) * 128 + (r & 7) * 16 + (k & 3) * 4; }

__device__ inline uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);                    // start address, 16-byte units
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;                 // leading byte offset: next core matrix along K
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;                 // stride byte offset: next core matrix along M / N
    d |= (uint64_t)1 << 46;                                      // descriptor version of sm_100
    return d;                                                    // layout type 0 = no swizzle
}
// kind::tf32, fp32 accumulate, A and B K-major
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__device__ inline void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ inline void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

struct Timing { long long split, mma, ldtm, total; };

template <int SPLITS>      // 1: plain TF32, 3: 3xTF32
__global__ void __launch_bounds__(NT, 1) dft_gemm(const float* __restrict__ in, const float* __restrict__ Bhi, const float* __restrict__ Blo,
                                                  float* __restrict__ out, Timing* tim, int iters)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float* rowmajor = reinterpret_cast<float*>(smem);                          // [64][96] input, later output
    unsigned char* a_hi = smem + M * RS * 4;
    unsigned char* a_lo = a_hi + A_BYTES;
    unsigned char* b_hi = a_lo + A_BYTES;
    unsigned char* b_lo = b_hi + B_BYTES;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);

    for (int q = tid; q < M * K; q += NT) rowmajor[(q / K) * RS + q % K] = in[(size_t)blockIdx.x * M * K + q];
    for (int q = tid; q < N * K; q += NT) {                                     // B[n][k] = F[k][n], already split on the host
        const int n = q / K, k = q % K;
        *reinterpret_cast<float*>(b_hi + core_off(n, k, B_RG)) = Bhi[q];
        *reinterpret_cast<float*>(b_lo + core_off(n, k, B_RG)) = Blo[q];
    }
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(a_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(a_lo);
    const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(b_hi), b_lo_s = (uint32_t)__cvta_generic_to_shared(b_lo);

    long long t_split = 0, t_mma = 0, t_ldtm = 0, t_total = 0;
    uint32_t parity = 0;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        const long long c0 = clock64();
        // ---- split: row-major fp32 -> hi (tf32 = upper 19 bits) and lo = a - hi, in the operand layout ----
        for (int q = tid; q < M * K / 4; q += NT) {
            const int r = q % M, k4 = q / M;                                    // consecutive threads: consecutive rows of one K chunk
            const float4 v = *reinterpret_cast<const float4*>(rowmajor + r * RS + 4 * k4);
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
            const int off = core_off(r, 4 * k4, A_RG);
            *reinterpret_cast<float4*>(a_hi + off) = h;
            if (SPLITS == 3) *reinterpret_cast<float4*>(a_lo + off) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        const long long c1 = clock64();
        // ---- tensor core: one thread issues, everybody waits on the mbarrier the commit arrives on ----
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            uint32_t acc = 0;
#pragma unroll
            for (int s = 0; s < SPLITS; ++s) {
                const uint32_t as = (s == 1) ? a_lo_s : a_hi_s, bs = (s == 2) ? b_lo_s : b_hi_s;
#pragma unroll
                for (int k = 0; k < K / 8; ++k) {                               // two K chunks per instruction
                    mma_tf32(tmem, smem_desc(as + k * 2 * A_RG * 128, A_RG * 128, 128), smem_desc(bs + k * 2 * B_RG * 128, B_RG * 128, 128), acc);
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
        }
        mbar_wait(bar_a, parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;");
        const long long c2 = clock64();
        // ---- read back: M = 64 puts rows 16 w .. 16 w + 15 on lanes 0..15 of warp w's quarter of tensor memory ----
        if (warp < 4) {
#pragma unroll
            for (int c = 0; c < N; c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                               "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                               "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                               "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (lane < 16) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(rowmajor + (16 * warp + lane) * RS + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
        }
        __syncthreads();
        const long long c3 = clock64();
        t_split += c1 - c0; t_mma += c2 - c1; t_ldtm += c3 - c2; t_total += c3 - c0;
        if (it + 1 < iters) {                                                    // next iteration transforms the same input again
            for (int q = tid; q < M * K; q += NT) out[(size_t)blockIdx.x * M * K + q] = rowmajor[(q / K) * RS + q % K];
            __syncthreads();
            for (int q = tid; q < M * K; q += NT) rowmajor[(q / K) * RS + q % K] = in[(size_t)blockIdx.x * M * K + q];
        }
    }
    for (int q = tid; q < M * K; q += NT) out[(size_t)blockIdx.x * M * K + q] = rowmajor[(q / K) * RS + q % K];
    if (tid == 0) tim[blockIdx.x] = Timing{t_split / iters, t_mma / iters, t_ldtm / iters, t_total / iters};
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main()
{
    const int ctas = 148, iters = 200;
    // real DFT matrix: column m < 49: cos(2 pi k m / 96), column 48 + m (m = 1..47): -sin(2 pi k m / 96)
    std::vector<double> F((size_t)K * N);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            const int m = n < 49 ? n : n - 48;
            const double th = 2 * M_PI * ((k * m) % 96) / 96.0;
            F[(size_t)k * N + n] = n < 49 ? cos(th) : -sin(th);
        }
    std::vector<float> Bhi((size_t)N * K), Blo((size_t)N * K), in((size_t)ctas * M * K);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float f = (float)F[(size_t)k * N + n];
            uint32_t u; memcpy(&u, &f, 4); u &= 0xFFFFE000u;
            float h; memcpy(&h, &u, 4);
            Bhi[(size_t)n * K + k] = h; Blo[(size_t)n * K + k] = (float)(F[(size_t)k * N + n] - (double)h);
        }
    srand(1);
    for (auto& v : in) v = (float)rand() / RAND_MAX * 2 - 1;
    float *d_in, *d_out, *d_bh, *d_bl; Timing* d_t;
    CK(cudaMalloc(&d_in, in.size() * 4)); CK(cudaMalloc(&d_out, in.size() * 4));
    CK(cudaMalloc(&d_bh, Bhi.size() * 4)); CK(cudaMalloc(&d_bl, Blo.size() * 4)); CK(cudaMalloc(&d_t, ctas * sizeof(Timing)));
    CK(cudaMemcpy(d_in, in.data(), in.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bh, Bhi.data(), Bhi.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bl, Blo.data(), Blo.size() * 4, cudaMemcpyHostToDevice));
    const int smem = M * RS * 4 + 2 * A_BYTES + 2 * B_BYTES + 1024;
    CK(cudaFuncSetAttribute(dft_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(dft_gemm<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> out(in.size());
    std::vector<Timing> tim(ctas);
    for (int splits : {1, 3}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            if (splits == 1) dft_gemm<1><<<ctas, NT, smem>>>(d_in, d_bh, d_bl, d_out, d_t, iters);
            else dft_gemm<3><<<ctas, NT, smem>>>(d_in, d_bh, d_bl, d_out, d_t, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(tim.data(), d_t, ctas * sizeof(Timing), cudaMemcpyDeviceToHost));
        double err = 0, ref_max = 0;
        for (int c : {0, 73, 147})
            for (int r = 0; r < M; ++r)
                for (int n = 0; n < N; ++n) {
                    double acc = 0;
                    for (int k = 0; k < K; ++k) acc += (double)in[((size_t)c * M + r) * K + k] * F[(size_t)k * N + n];
                    err = fmax(err, fabs(acc - (double)out[((size_t)c * M + r) * K + n]));
                    ref_max = fmax(ref_max, fabs(acc));
                }
        double s = 0, m = 0, l = 0, t = 0;
        for (auto& x : tim) { s += x.split; m += x.mma; l += x.ldtm; t += x.total; }
        printf("{\"variant\": \"%dxTF32\", \"mma_instructions\": %d, \"cycles_split\": %.0f, \"cycles_mma\": %.0f, \"cycles_ldtm\": %.0f, \"cycles_total\": %.0f, "
               "\"max_abs_err\": %.3e, \"max_abs_ref\": %.2f, \"rel_err\": %.2e, \"launch_ms\": %.3f}\n",
               splits, 12 * splits, s / ctas, m / ctas, l / ctas, t / ctas, err, ref_max, err / ref_max, ms);
    }
    return 0;
}
