// Microbenchmark: the tendency march of the 96 x 64 kernel with two thread mappings.
//   A: the product code (rbc2d_core.h phase_edge_fluxes + phase_tendency): 384 threads, one column per lane, x-fluxes
//      shared through shuffles and an edge table;
//   B: three adjacent columns per lane: 96 = 3 x 32 puts one periodic row on one warp (right fluxes of the last column by
//      a periodic shuffle, the other faces shared inside the thread, x-windows of 8 values for 3 cells); 256 threads =
//      32 lanes x 8 strips of 8 rows, up to 255 registers.
// Both run `iters` stages (tendency + buffer swap, no projection, tiny dt) on one environment per CTA held in shared
// memory and must produce the same state.  Decides whether mapping B is worth re-tiling the whole kernel for.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rbc_gym_b200/csrc -o tools/microbench/tendency_tile tools/microbench/tendency_tile.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "rbc2d_core.h"

using namespace rbc2d;

constexpr int SMEM_WORDS = 2 * NS_SM + NX + NE;

// ------------------------------------------------------------------------------------------------
// mapping B
// ------------------------------------------------------------------------------------------------
constexpr int BC = 3, BSTRIP = 8, BRS = NZ / BSTRIP, BNT = 32 * BSTRIP;

__device__ __forceinline__ float next_lane(float v) { return __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31); }

__device__ void tendency_b(int tid, const Consts<float>& C, const float* __restrict__ c, float* __restrict__ n, const float* __restrict__ Tb,
                           const float* gm_in, float* gm_out, float dt, float gam, float zet, bool use_gm)
{
    const int lane = tid & 31, s = tid >> 5, k0 = s * BRS, c0 = BC * lane;
    const float* __restrict__ cb = c + OFF_B;
    const float* __restrict__ cu = c + OFF_U;
    const float* __restrict__ cw = c + OFF_W;
    int xl[5];                                            // the 5 columns of the 8-wide x-window that are not owned
    xl[0] = wrapx(c0 - 3); xl[1] = wrapx(c0 - 2); xl[2] = wrapx(c0 - 1); xl[3] = wrapx(c0 + 3); xl[4] = wrapx(c0 + 4);

    float bz[BC][7], uz[BC][7], wz[BC][7], wr[8];
#pragma unroll
    for (int q = 0; q < BC; ++q)
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int k = k0 - 3 + j;
            const bool ok = (k >= 0 && k < NZ);
            bz[q][j] = ok ? cb[k * SX + c0 + q] : 0.f;
            uz[q][j] = ok ? cu[k * SX + c0 + q] : 0.f;
            wz[q][j] = (k >= 0 && k <= NZ) ? cw[k * SX + c0 + q] : 0.f;
        }
    wr[0] = cw[k0 * SX + xl[0]]; wr[1] = cw[k0 * SX + xl[1]]; wr[2] = cw[k0 * SX + xl[2]];
    wr[3] = wz[0][3]; wr[4] = wz[1][3]; wr[5] = wz[2][3];
    wr[6] = cw[k0 * SX + xl[3]]; wr[7] = cw[k0 * SX + xl[4]];

    float Fzb_lo[BC], Wu_lo[BC], Ww_lo[BC];
#pragma unroll
    for (int q = 0; q < BC; ++q) {
        Fzb_lo[q] = 0.f; Wu_lo[q] = 0.f; Ww_lo[q] = 0.f;
        if (k0 >= 1) {
            Fzb_lo[q] = upwind5(wz[q][3], bz[q]);
            Wu_lo[q] = upwind5(centred4(wr[q + 1], wr[q + 2], wr[q + 3], wr[q + 4]), uz[q]);
            Ww_lo[q] = upwind5(centred4(wz[q][1], wz[q][2], wz[q][3], wz[q][4]), wz[q]);
        }
    }
    float tb[BC];
#pragma unroll
    for (int q = 0; q < BC; ++q) tb[q] = Tb[c0 + q];
    const float kdx = C.kappa * C.idx2, kdz = C.kappa * C.idz2, ndx = C.nu * C.idx2, ndz = C.nu * C.idz2;
    const float dtg = dt * gam, dtz = dt * zet;

    auto row = [&](auto edge_tag, const int r) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int k = k0 + r;
        float g0[3][BC];
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
            for (int q = 0; q < BC; ++q) g0[f][q] = use_gm ? gm_in[((f * BC + q) * BRS + r) * BNT + tid] : 0.f;
        float xb[8], xu[8], xw[8];
        xb[0] = cb[k * SX + xl[0]]; xb[1] = cb[k * SX + xl[1]]; xb[2] = cb[k * SX + xl[2]];
        xb[3] = bz[0][3]; xb[4] = bz[1][3]; xb[5] = bz[2][3];
        xb[6] = cb[k * SX + xl[3]]; xb[7] = cb[k * SX + xl[4]];
        xu[0] = cu[k * SX + xl[0]]; xu[1] = cu[k * SX + xl[1]]; xu[2] = cu[k * SX + xl[2]];
        xu[3] = uz[0][3]; xu[4] = uz[1][3]; xu[5] = uz[2][3];
        xu[6] = cu[k * SX + xl[3]]; xu[7] = cu[k * SX + xl[4]];
        xw[0] = cw[(k + 1) * SX + xl[0]]; xw[1] = cw[(k + 1) * SX + xl[1]]; xw[2] = cw[(k + 1) * SX + xl[2]];
        xw[3] = wz[0][4]; xw[4] = wz[1][4]; xw[5] = wz[2][4];
        xw[6] = cw[(k + 1) * SX + xl[3]]; xw[7] = cw[(k + 1) * SX + xl[4]];
        const bool top = EDGE && (k == NZ - 1);
        const bool bot = EDGE && (k == 0);
        const int o_face_hi = EDGE ? ord_up_face(k + 1) : 5;
        const int o_ce_face = EDGE ? ord_ce_face(k) : 4;
        const int o_up_cen = EDGE ? ord_up_cen(k) : 5;
        const int o_ce_cen = EDGE ? ord_ce_cen(k) : 4;

        // left x-fluxes of the three columns, then the right flux of the last one from the next lane
        float Fx[BC + 1], Fu[BC + 1], Fw[BC + 1];
#pragma unroll
        for (int q = 0; q < BC; ++q) {
            Fx[q] = upwind5(xu[q + 3], xb + q);
            Fu[q] = upwind5(centred4(xu[q + 1], xu[q + 2], xu[q + 3], xu[q + 4]), xu + q);
            Fw[q] = upwind5(centred_ord(uz[q][1], uz[q][2], uz[q][3], uz[q][4], o_ce_face), wr + q);
        }
        Fx[BC] = next_lane(Fx[0]); Fu[BC] = next_lane(Fu[0]); Fw[BC] = next_lane(Fw[0]);
#pragma unroll
        for (int q = 0; q < BC; ++q) {
            const float Fzb_hi = top ? 0.f : upwind_ord(wz[q][4], bz[q] + 1, o_face_hi);
            const float Wu_hi = top ? 0.f : upwind_ord(centred4(xw[q + 1], xw[q + 2], xw[q + 3], xw[q + 4]), uz[q] + 1, o_face_hi);
            const float Ww_hi = upwind_ord(centred_ord(wz[q][2], wz[q][3], wz[q][4], wz[q][5], o_ce_cen), wz[q] + 1, o_up_cen);
            const float bdn = bot ? (2.f * tb[q] - bz[q][3]) : bz[q][2];
            const float bup = top ? (2.f * C.b_top - bz[q][3]) : bz[q][4];
            const float Gb = (Fx[q] - Fx[q + 1]) * C.idx + (Fzb_lo[q] - Fzb_hi) * C.idz + (xb[q + 4] - 2.f * xb[q + 3] + xb[q + 2]) * kdx +
                             (bup - 2.f * bz[q][3] + bdn) * kdz;
            const float udn = bot ? -uz[q][3] : uz[q][2];
            const float uup = top ? -uz[q][3] : uz[q][4];
            const float Gu = (Fu[q] - Fu[q + 1]) * C.idx + (Wu_lo[q] - Wu_hi) * C.idz + (xu[q + 4] - 2.f * xu[q + 3] + xu[q + 2]) * ndx +
                             (uup - 2.f * uz[q][3] + udn) * ndz;
            float Gw = (Fw[q] - Fw[q + 1]) * C.idx + (Ww_lo[q] - Ww_hi) * C.idz + (wr[q + 4] - 2.f * wr[q + 3] + wr[q + 2]) * ndx +
                       (wz[q][4] - 2.f * wz[q][3] + wz[q][2]) * ndz;
            Gw += 0.5f * (bz[q][2] + bz[q][3]);
            if (bot) Gw = 0.f;
            gm_out[((0 * BC + q) * BRS + r) * BNT + tid] = Gb;
            gm_out[((1 * BC + q) * BRS + r) * BNT + tid] = Gu;
            gm_out[((2 * BC + q) * BRS + r) * BNT + tid] = Gw;
            n[OFF_B + k * SX + c0 + q] = bz[q][3] + dtg * Gb + dtz * g0[0][q];
            n[OFF_U + k * SX + c0 + q] = uz[q][3] + dtg * Gu + dtz * g0[1][q];
            n[OFF_W + k * SX + c0 + q] = bot ? 0.f : wz[q][3] + dtg * Gw + dtz * g0[2][q];
            Fzb_lo[q] = Fzb_hi; Wu_lo[q] = Wu_hi; Ww_lo[q] = Ww_hi;
#pragma unroll
            for (int j = 0; j < 6; ++j) { bz[q][j] = bz[q][j + 1]; uz[q][j] = uz[q][j + 1]; wz[q][j] = wz[q][j + 1]; }
            const int kn = k + 4;
            bz[q][6] = (kn < NZ) ? cb[kn * SX + c0 + q] : 0.f;
            uz[q][6] = (kn < NZ) ? cu[kn * SX + c0 + q] : 0.f;
            wz[q][6] = (kn <= NZ) ? cw[kn * SX + c0 + q] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[j] = xw[j];
    };

#pragma unroll
    for (int r = 0; r < BRS; ++r) {
        if (r >= 2 && r < BRS - 3) {
            row(BoolTag<false>{}, r);
        } else {
            const bool edge = (r < 2) ? (s == 0) : (s == BSTRIP - 1);
            if (edge) row(BoolTag<true>{}, r);
            else row(BoolTag<false>{}, r);
        }
    }
    if (s == BSTRIP - 1)
        for (int q = 0; q < BC; ++q) n[OFF_W + NZ * SX + c0 + q] = 0.f;
}

// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(MODE == 0 ? NT : BNT, 1)
march(Consts<float> C, const float* state0, float* out, float* gm_all, int iters, float dt)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int THREADS = MODE == 0 ? NT : BNT;
    float* s0 = sm;
    float* s1 = sm + NS_SM;
    float* Tb = sm + 2 * NS_SM;
    float* E = Tb + NX;
    float* gm = gm_all + (size_t)blockIdx.x * 2 * NSTATE;
    for (int q = threadIdx.x; q < NS_SM; q += THREADS) { s0[q] = 0.f; s1[q] = 0.f; }
    __syncthreads();
    for (int q = threadIdx.x; q < NSTATE; q += THREADS) {
        const int f = q < NCELL ? 0 : (q < 2 * NCELL ? 1 : 2);
        const int rem = q - (f == 0 ? 0 : (f == 1 ? NCELL : 2 * NCELL));
        s0[(f == 0 ? OFF_B : (f == 1 ? OFF_U : OFF_W)) + (rem / NX) * SX + rem % NX] = state0[q];
    }
    if (threadIdx.x < NX) Tb[threadIdx.x] = 2.f + 0.5f * sinf(0.3f * threadIdx.x);
    __syncthreads();
    float* cur = s0;
    float* nxt = s1;
    const float gam[3] = {8.f / 15.f, 5.f / 12.f, 3.f / 4.f}, zet[3] = {0.f, -17.f / 60.f, -5.f / 12.f};
    for (int it = 0; it < iters; ++it) {
        const int stage = it % 3;
        const float* gin = gm + ((stage & 1) ? 0 : NSTATE);
        float* gout = gm + ((stage & 1) ? NSTATE : 0);
        if (MODE == 0) {
            phase_edge_fluxes(threadIdx.x, cur, E);
            __syncthreads();
            phase_tendency<float, false>(threadIdx.x, C, cur, nxt, nullptr, Tb, E, gin, gout, dt, gam[stage], zet[stage], stage > 0);
        } else {
            tendency_b(threadIdx.x, C, cur, nxt, Tb, gin, gout, dt, gam[stage], zet[stage], stage > 0);
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    float* o = out + (size_t)blockIdx.x * NSTATE;
    for (int q = threadIdx.x; q < NSTATE; q += THREADS) {
        const int f = q < NCELL ? 0 : (q < 2 * NCELL ? 1 : 2);
        const int rem = q - (f == 0 ? 0 : (f == 1 ? NCELL : 2 * NCELL));
        o[q] = cur[(f == 0 ? OFF_B : (f == 1 ? OFF_U : OFF_W)) + (rem / NX) * SX + rem % NX];
    }
}

int main()
{
    HostConfig h{1e5, 0.7, 2 * 3.14159265358979323846, 2.0, 1.0, 0.75, 1.0, 0.03, 300.0, 12, 8, 48, 3};
    Consts<float> C = make_consts<float>(h);
    const int ctas = 148, iters = 3000;
    std::vector<float> st(NSTATE);
    for (int q = 0; q < NSTATE; ++q) {
        const int f = q < NCELL ? 0 : (q < 2 * NCELL ? 1 : 2);
        const int rem = q - (f == 0 ? 0 : (f == 1 ? NCELL : 2 * NCELL));
        const int k = rem / NX, i = rem % NX;
        const double x = 2 * 3.14159265358979323846 * i / NX, z = (k + 0.5) / NZ;
        if (f == 0) st[q] = (float)(2.0 - z + 0.1 * sin(3 * x) * sin(3.14159 * z));
        else if (f == 1) st[q] = (float)(0.3 * sin(2 * x + 0.3) * cos(3.14159 * z));
        else st[q] = (k == 0 || k == NZ) ? 0.f : (float)(0.3 * cos(2 * x) * sin(3.14159 * k / NZ));
    }
    float *d_st, *d_out[2], *d_gm;
    cudaMalloc(&d_st, NSTATE * 4);
    cudaMemcpy(d_st, st.data(), NSTATE * 4, cudaMemcpyHostToDevice);
    for (int m = 0; m < 2; ++m) cudaMalloc(&d_out[m], (size_t)ctas * NSTATE * 4);
    cudaMalloc(&d_gm, (size_t)ctas * 2 * NSTATE * 4);
    const size_t smem = SMEM_WORDS * 4;
    cudaFuncSetAttribute(march<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(march<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(d_gm, 0, (size_t)ctas * 2 * NSTATE * 4);
            cudaEventRecord(e0);
            if (mode == 0) march<0><<<ctas, NT, smem>>>(C, d_st, d_out[0], d_gm, iters, 1e-4f);
            else march<1><<<ctas, BNT, smem>>>(C, d_st, d_out[1], d_gm, iters, 1e-4f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%s: %.3f ms  %.0f cycles per tendency stage (1.965 GHz)  %s\n", mode ? "B (3 columns per lane, 256 thr)" : "A (product, 384 thr)          ",
                   ms, ms * 1e-3 * 1.965e9 / iters, cudaGetErrorString(cudaGetLastError()));
        }
    std::vector<float> a(NSTATE), b(NSTATE);
    cudaMemcpy(a.data(), d_out[0], NSTATE * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), d_out[1] + (size_t)77 * NSTATE, NSTATE * 4, cudaMemcpyDeviceToHost);
    double md = 0, mx = 0;
    for (int q = 0; q < NSTATE; ++q) { md = fmax(md, fabs((double)a[q] - b[q])); mx = fmax(mx, fabs((double)a[q])); }
    printf("max |A - B| = %.3e  (max |A| = %.3f)\n", md, mx);
    return 0;
}
