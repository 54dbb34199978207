// rbc3dg_lib.cu — sm_100a kernels of the stage-streaming 3D action step for arbitrary grids (rbc3dg_core.h) and the plan
// object rbc3d_lib.cu drives them through.  Every kernel runs over the whole batch (blockIdx.y = environment), so a single
// 64 x 64 x 32 environment already fills the GPU with 131 072 threads per launch.
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "rbc3dg_api.h"
#include "rbc3dg_core.h"
#include "rbc_common.h"

using namespace rbc3dg;

namespace {

constexpr int TB = 256;      // threads per block of the per-cell kernels

template <typename Real>
__global__ void g3_heater_kernel(Dims D, ConstsG<Real> C, const float* actions, Real* Tb, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D.ncol) return;
    Tb[(size_t)env * D.ncol + c] = (Real)heater_patch_T(C.heaters, C.heater_limit, C.b_hot, actions + (size_t)env * C.heaters * C.heaters,
                                                        c % D.nx, c / D.nx, D.nx, D.ny);
}

template <typename Real>
__global__ void __launch_bounds__(TB)
g3_tendency_kernel(Dims D, ConstsG<Real> C, const Real* nu_env, const Real* kappa_env, const Real* S_all, Real* P_all, Real* G_all,
                   const Real* Tb_all, const int* env_ids, Real dt, Real gam, Real zet, int use_prev, int store_g)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= D.nc) return;
    cell_tendency<Real>(D, C, nu_env[env], kappa_env[env], S_all + (size_t)env * D.nstate, P_all + (size_t)env * D.nstate,
                        G_all + (size_t)env * 4 * D.nc, Tb_all + (size_t)env * D.ncol, cell, dt, gam, zet, use_prev != 0, store_g != 0);
}

// ------------------------------------------------------------------------------------------
// Tiled tendency (grids with nx % 32 == 0, ny % 8 == 0).  A CTA owns a 32 x 8 (or 32 x 16) patch of columns and marches it up through
// the levels: one thread per column, the z-windows of its own column slide through registers (one coalesced global load per field
// and level), and everything it needs from its neighbours comes from a RING OF LEVEL PLANES WITH HALO in shared memory
// (40 x 14 values per field: 4 + 32 + 4 columns, 3 + 8 + 3 rows; the periodic wrap is resolved once, when a thread computes which
// 16-byte chunks of a level it copies), so every stencil access of the march is a shared-memory load at a compile-time offset from
// the thread's own slot.  Level k reads planes k and k + 1; plane k + 5 is copied meanwhile (cp.async, 16 bytes per instruction, no
// register staging); the planes are handed over through mbarriers, not CTA barriers (see PIPE_R below).  Same arithmetic as the
// per-cell kernel (tendency_from_windows), 1 480 -> 685 instructions per cell.
// ------------------------------------------------------------------------------------------
constexpr int TT_X = 32, TT_PW = 40;                   // patch width; plane row = 4 + 32 + 4 columns (3 + 32 + 3 read)
template <int TT_Y>
struct TileGeom {
    static constexpr int NT = TT_X * TT_Y;             // threads
    static constexpr int PH = TT_Y + 6;                // plane rows
    static constexpr int PLANE = TT_PW * PH, SLOT = 4 * PLANE;
};

template <typename Real>
__device__ __forceinline__ void async_copy_chunk(Real* dst_shared, const Real* src_global)      // 16 bytes, both 16-byte aligned
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global) : "memory");
}

// No CTA barrier in the level loop: the ring has PIPE_R planes guarded by one "full" and one "empty" mbarrier per slot.  A thread's
// copies of a plane arrive on the slot's full barrier when they land (cp.async.mbarrier.arrive), every thread arrives on the empty
// barrier of a plane when it has finished the level that reads it last, and the copy of plane k + PIPE_DIST waits for the empty barrier
// of the plane it overwrites (PIPE_R - PIPE_DIST = 3 levels back): the warps of a CTA may drift a few levels apart instead of meeting
// at every level.  Measured against the form with one __syncthreads per level (ring of four): 1378 against 1298 env-steps/s at
// (R, DIST) = (6, 4); (5, 3) 1309, (7, 4) 1378, (7, 5) 1376, (8, 5) 1400.
// fp32: 128 registers per thread, i.e. 512 resident threads per SM — one 16-row CTA or two 8-row CTAs; fp64 needs ~240 registers and
// runs one CTA per SM
constexpr int PIPE_R = 8, PIPE_DIST = 5;
__device__ __forceinline__ void pipe_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void pipe_mbar_wait(unsigned bar, unsigned phase)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    } while (!ok);
}
template <typename Real, int TT_Y>
__global__ void __launch_bounds__(TT_X * TT_Y, sizeof(Real) == 4 ? 512 / (TT_X * TT_Y) : 1)
g3_tendency_tiled_kernel(Dims D, ConstsG<Real> C, const Real* nu_env, const Real* kappa_env, const Real* S_all, Real* P_all, Real* G_all,
                         const Real* Tb_all, const int* env_ids, Real dt, Real gam, Real zet, int use_prev, int store_g)
{
    constexpr int TT_NT = TileGeom<TT_Y>::NT, TT_PLANE = TileGeom<TT_Y>::PLANE, TT_SLOT = TileGeom<TT_Y>::SLOT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Real* ring = reinterpret_cast<Real*>(smem_raw);
    constexpr int R = PIPE_R;
    const unsigned bars = (unsigned)__cvta_generic_to_shared(smem_raw + (size_t)R * TT_SLOT * sizeof(Real));      // full[R], empty[R]
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const int nx = D.nx, ny = D.ny, nz = D.nz, ncol = D.ncol;
    const int tiles_x = nx / TT_X;
    const int x0 = (blockIdx.x % tiles_x) * TT_X, y0 = (blockIdx.x / tiles_x) * TT_Y;
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    const int i = x0 + lx, j = y0 + ly, colz = j * nx + i;
    const Real* __restrict__ S = S_all + (size_t)env * D.nstate;
    Real* P = P_all + (size_t)env * D.nstate;
    Real* G = G_all + (size_t)env * 4 * D.nc;
    const Real nu = nu_env[env], kappa = kappa_env[env];
    const Real tb = Tb_all[(size_t)env * ncol + colz];

    // which 16-byte chunks of a level this thread copies.  A plane row starts four columns left of the patch (x0 is a multiple of 32,
    // so every chunk is 16-byte aligned in global and in shared memory, and the periodic wrap moves whole chunks): TT_PW = 40 values
    // = 4 + 32 + 4, of which the march reads 3 + 32 + 3.  All four fields of a level are one list of 4 * PH * (TT_PW / CH) chunks
    // (560 per level for 8-row fp32 patches: 3 copy instructions per thread where the element-wise version issued 12).
    constexpr int CH = 16 / (int)sizeof(Real), CPR = TT_PW / CH, NCHF = TileGeom<TT_Y>::PH * CPR, NCH = 4 * NCHF, NQ = (NCH + TT_NT - 1) / TT_NT;
    int goff[NQ], soff[NQ];      // element offsets inside the state (field offset included) and inside a ring slot; goff < 0: nothing to copy
    bool isw[NQ];                // chunk of w (copied for level nz too)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int e = tid + q * TT_NT;
        const int f = e / NCHF, r = e % NCHF, pr = r / CPR, pc = r % CPR;
        goff[q] = e < NCH ? (f * D.nc + ((y0 - 3 + pr) & (ny - 1)) * nx + ((x0 - 4 + pc * CH) & (nx - 1))) : -1;
        soff[q] = f * TT_PLANE + pr * TT_PW + pc * CH;
        isw[q] = f == 3;
    }
    auto copy_plane = [&](int level) {                        // level <= nz: w has a face there, the cell-centred fields do not
        const int sl = level % R;
        Real* slot = ring + sl * TT_SLOT;
        const Real* src = S + level * ncol;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (goff[q] >= 0 && (level < nz || isw[q])) async_copy_chunk(slot + soff[q], src + goff[q]);
        // the thread's arrival on the plane's "full" barrier happens when its copies have landed
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bars + 8u * sl) : "memory");
    };
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 2 * R; ++q) pipe_mbar_init(bars + 8u * q, TT_NT);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < PIPE_DIST; ++l) copy_plane(l);          // nz >= 6 > PIPE_DIST - 1

    // z-windows of the own column for level 0 (slot d <-> level d - 3, clamped like the per-cell kernel)
    Windows<Real> W;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        const int kc = d - 3 < 0 ? 0 : (d - 3 > nz - 1 ? nz - 1 : d - 3), kw = d - 3 < 0 ? 0 : (d - 3 > nz ? nz : d - 3);
        W.bz[d] = S[D.gb + kc * ncol + colz]; W.uz[d] = S[D.gu + kc * ncol + colz]; W.vz[d] = S[D.gv + kc * ncol + colz];
        W.wz[d] = S[D.gw + kw * ncol + colz];
    }
    pipe_mbar_wait(bars, 0);                                    // plane 0 is in (plane 1 is awaited by level 0)
    const int p0 = (ly + 3) * TT_PW + lx + 4;
    {   // neighbour-column z-windows: slots 0..2 <-> levels -2, -1, 0 (all level 0 after the clamp; slot 3 is filled per level)
        const Real* s0 = ring;
        const Real un = s0[TT_PLANE + p0 + 1], vn = s0[2 * TT_PLANE + p0 + TT_PW];
        W.u_ip_z[0] = un; W.u_ip_z[1] = un; W.u_ip_z[2] = un; W.u_ip_z[3] = un;
        W.v_jp_z[0] = vn; W.v_jp_z[1] = vn; W.v_jp_z[2] = vn; W.v_jp_z[3] = vn;
    }

    const Real* Scol = S + colz;                             // the thread's own column: field b, level 0 of the state, the predicted state
    Real* Pcol = P + colz;                                  // and the tendency slab (fields nc values apart)
    Real* Gcol = G + colz;
    Tend<Real> zf{Real(0), Real(0), Real(0), Real(0)};      // fluxes through the lower face of the current level (wall: zero)
    for (int k = 0; k < nz; ++k) {
        if (k + PIPE_DIST <= nz) {
            // the slot of plane k + DIST held plane k + DIST - R, last read at that level: wait until everybody has released it
            const int old = k + PIPE_DIST - R;
            if (old >= 0) pipe_mbar_wait(bars + 8u * (R + old % R), (unsigned)((old / R) & 1));
            copy_plane(k + PIPE_DIST);
        }
        pipe_mbar_wait(bars + 8u * ((k + 1) % R), (unsigned)(((k + 1) / R) & 1));      // plane k + 1 (plane k was awaited a level ago)
        Real* Gk = Gcol + k * ncol;
        const Tend<Real> prev = load_prev_at<Real>(Gk, D.nc, use_prev != 0);
        const Real* s0 = ring + (k % R) * TT_SLOT;
        const Real* s1 = ring + ((k + 1) % R) * TT_SLOT;
#pragma unroll
        for (int d = 0; d < 7; ++d) {
            W.bx[d] = s0[p0 + d - 3]; W.by[d] = s0[p0 + (d - 3) * TT_PW];
            W.ux[d] = s0[TT_PLANE + p0 + d - 3]; W.uy[d] = s0[TT_PLANE + p0 + (d - 3) * TT_PW];
            W.vx[d] = s0[2 * TT_PLANE + p0 + d - 3]; W.vy[d] = s0[2 * TT_PLANE + p0 + (d - 3) * TT_PW];
            W.wx[d] = s0[3 * TT_PLANE + p0 + d - 3]; W.wy[d] = s0[3 * TT_PLANE + p0 + (d - 3) * TT_PW];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            W.v_jp[q] = s0[2 * TT_PLANE + p0 + TT_PW + q - 2];
            W.u_ip[q] = s0[TT_PLANE + p0 + 1 + (q - 2) * TT_PW];
            W.w_kp_x[q] = s1[3 * TT_PLANE + p0 + q - 2];
            W.w_kp_y[q] = s1[3 * TT_PLANE + p0 + (q - 2) * TT_PW];
        }
        if (k + 1 < nz) { W.u_ip_z[3] = s1[TT_PLANE + p0 + 1]; W.v_jp_z[3] = s1[2 * TT_PLANE + p0 + TT_PW]; }
        const Tend<Real> g = interior_level(k, nz) ? tendency_from_windows_t<Real, true, true>(C, nu, kappa, nz, k, W, tb, &zf)
                                                   : tendency_from_windows_t<Real, false, true>(C, nu, kappa, nz, k, W, tb, &zf);
        rk3_substep_store_at<Real>(D, Pcol + k * ncol, Gk, k, W, g, prev, dt, gam, zet, store_g != 0);
        // slide the register windows up one level
#pragma unroll
        for (int d = 0; d < 6; ++d) { W.bz[d] = W.bz[d + 1]; W.uz[d] = W.uz[d + 1]; W.vz[d] = W.vz[d + 1]; W.wz[d] = W.wz[d + 1]; }
        {
            const int kn = k + 4, kc = kn > nz - 1 ? nz - 1 : kn, kw = kn > nz ? nz : kn;
            const Real* Sk = Scol + kc * ncol;
            W.bz[6] = Sk[0]; W.uz[6] = Sk[D.nc]; W.vz[6] = Sk[2 * D.nc];
            W.wz[6] = Scol[3 * D.nc + kw * ncol];
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) { W.u_ip_z[q] = W.u_ip_z[q + 1]; W.v_jp_z[q] = W.v_jp_z[q + 1]; }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8u * (R + k % R)) : "memory");      // done with plane k
    }
}

// one FFT stage on the device: the CTA's threads stride over the butterflies, then a barrier
struct BlockRun {
    template <typename F>
    __device__ void operator()(int n_items, F fn) const
    {
        for (int it = threadIdx.x; it < n_items; it += blockDim.x) fn(it);
        __syncthreads();
    }
};
// the same for a CTA size and (after inlining, with compile-time plane extents) an item count known to the compiler: the item loop
// unrolls and the index arithmetic of every butterfly folds into constants
template <int NT>
struct BlockRunFixed {
    template <typename F>
    __device__ __forceinline__ void operator()(int n_items, F fn) const
    {
#pragma unroll
        for (int it = threadIdx.x; it < n_items; it += NT) fn(it);
        __syncthreads();
    }
};
constexpr int FFT_NT = 256;      // CTA size of the plane kernels instantiated with compile-time extents
// Plane extents known at compile time (LX2 = log2 nx, LY2 = log2 ny; 0 = the run-time grid).  Two thirds of the ~130 instructions of a
// radix-4 step of the run-time version are index arithmetic (run-time shifts, masks and strides); with constant extents it folds away.
// Same butterflies in the same order on the same values: bit-identical results.
template <int LX2, int LY2>
__device__ __forceinline__ Dims plane_dims(const Dims& D)
{
    if (LX2 == 0) return D;
    Dims F = D;
    F.nx = 1 << LX2; F.ny = 1 << LY2; F.ncol = 1 << (LX2 + LY2); F.lx2 = LX2; F.ly2 = LY2;
    F.nc = F.ncol * D.nz; F.nw = F.ncol * (D.nz + 1); F.nstate = 3 * F.nc + F.nw;
    F.gb = 0; F.gu = F.nc; F.gv = 2 * F.nc; F.gw = 3 * F.nc;
    return F;
}
// twiddles of both directions into shared memory, behind the plane: [nx/2] then [ny/2] complex
template <typename Real>
__device__ void stage_twiddles(const Dims& D, cx<Real>* dst, const cx<Real>* twx, const cx<Real>* twy)
{
    const int hx = D.nx >> 1, hy = D.ny >> 1;
    for (int q = threadIdx.x; q < hx + hy; q += blockDim.x) dst[q] = q < hx ? twx[q] : twy[q - hx];
}

// four consecutive values (16-byte aligned: rows are multiples of 8 values, field and environment strides multiples of 64) in one
// or two vector accesses.  The vectorised kernels below evaluate exactly the expressions of cell_divergence / cell_correct on the
// same values, four cells per thread, so that each thread has 64 ... 176 bytes in flight instead of 4 ... 48.
template <typename Real>
struct Quad {
    Real v[4];
};
__device__ __forceinline__ Quad<float> ld4(const float* p)
{
    const float4 t = *reinterpret_cast<const float4*>(p);
    return Quad<float>{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ Quad<double> ld4(const double* p)
{
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    return Quad<double>{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void st4(float* p, const Quad<float>& q) { *reinterpret_cast<float4*>(p) = make_float4(q.v[0], q.v[1], q.v[2], q.v[3]); }
__device__ __forceinline__ void st4(double* p, const Quad<double>& q)
{
    *reinterpret_cast<double2*>(p) = make_double2(q.v[0], q.v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(q.v[2], q.v[3]);
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<size_t>(p) & 15) == 0; }

// divergence of four cells of level k starting at column i (multiple of 4) of row j: the expression of cell_divergence
template <typename Real>
__device__ __forceinline__ Quad<Real> quad_divergence(const Dims& D, const ConstsG<Real>& C, const Real* P, int i, int j, int k, const Quad<Real>& wk,
                                                      const Quad<Real>& wk1)
{
    const int row = (k * D.ny + j) * D.nx, rowp = (k * D.ny + ((j + 1) & (D.ny - 1))) * D.nx;
    const Quad<Real> u = ld4(P + D.gu + row + i), v = ld4(P + D.gv + row + i), vp = ld4(P + D.gv + rowp + i);
    const Real ur = P[D.gu + row + ((i + 4) & (D.nx - 1))];
    Quad<Real> d;
#pragma unroll
    for (int m = 0; m < 4; ++m) d.v[m] = ((m < 3 ? u.v[m + 1] : ur) - u.v[m]) * C.idx + (vp.v[m] - v.v[m]) * C.idy + (wk1.v[m] - wk.v[m]) * C.idz;
    return d;
}

// divergence of TWO levels (2p in the real, 2p + 1 in the imaginary part, see mode_pair_thomas) into a shared-memory plane, then the
// forward FFT in x and y (decimation in frequency)
template <typename Real, int LX2 = 0, int LY2 = 0>
__global__ void __launch_bounds__(LX2 ? FFT_NT : 1024)
g3_div_fft_kernel(Dims D_arg, ConstsG<Real> C, const Real* P_all, cx<Real>* Z_all, const cx<Real>* twx, const cx<Real>* twy, const int* env_ids)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<Real>* Z = reinterpret_cast<cx<Real>*>(smem_raw);
    const Dims D = plane_dims<LX2, LY2>(D_arg);
    const int pitch = D.nx + 1;      // odd row pitch: see plane_fft_forward
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y, pz = blockIdx.x, nzp = (D.nz + 1) >> 1;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const Real* P = P_all + (size_t)env * D.nstate;
    const bool odd = 2 * pz + 1 < D.nz;
    if (aligned16(P)) {
        // four cells per thread and iteration; the w face between the two levels is loaded once
#pragma unroll 2
        for (int c = 4 * threadIdx.x; c < D.ncol; c += 4 * blockDim.x) {
            const int i = c & (D.nx - 1), j = c >> D.lx2, k = 2 * pz;
            const Quad<Real> w0 = ld4(P + D.gw + k * D.ncol + c), w1 = ld4(P + D.gw + (k + 1) * D.ncol + c);
            const Quad<Real> d0 = quad_divergence<Real>(D, C, P, i, j, k, w0, w1);
            Quad<Real> d1{{Real(0), Real(0), Real(0), Real(0)}};
            if (odd) d1 = quad_divergence<Real>(D, C, P, i, j, k + 1, w1, ld4(P + D.gw + (k + 2) * D.ncol + c));
#pragma unroll
            for (int m = 0; m < 4; ++m) Z[j * pitch + i + m] = cx<Real>{d0.v[m], d1.v[m]};
        }
    } else {
#pragma unroll 4
        for (int c = threadIdx.x; c < D.ncol; c += blockDim.x) {
            const int i = c & (D.nx - 1), j = c >> D.lx2;
            Z[j * pitch + i] = cx<Real>{cell_divergence<Real>(D, C, P, i, j, 2 * pz), odd ? cell_divergence<Real>(D, C, P, i, j, 2 * pz + 1) : Real(0)};
        }
    }
    cx<Real>* tw = Z + D.ny * pitch;
    stage_twiddles<Real>(D, tw, twx, twy);
    __syncthreads();
    if (LX2) plane_fft_forward<Real>(D, Z, tw, tw + (D.nx >> 1), BlockRunFixed<FFT_NT>{}, pitch, true);
    else plane_fft_forward<Real>(D, Z, tw, tw + (D.nx >> 1), BlockRun{}, pitch, true);
    Real* out = reinterpret_cast<Real*>(Z_all + ((size_t)env * nzp + pz) * D.ncol);      // Z_all comes from cudaMalloc, planes are multiples of 512 bytes
    for (int c = 2 * threadIdx.x; c < D.ncol; c += 2 * blockDim.x) {                         // two complex values = 16 bytes per store
        const int zi = (c >> D.lx2) * pitch + (c & (D.nx - 1));
        const cx<Real> a = Z[zi], b = Z[zi + 1];
        st4(out + 2 * c, Quad<Real>{{a.re, a.im, b.re, b.im}});
    }
}

template <typename Real>
__global__ void g3_thomas_kernel(Dims D, cx<Real>* Z_all, const Real* cp, Real scale, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= D.ncol) return;
    mode_pair_thomas<Real>(D, Z_all + (size_t)env * ((D.nz + 1) >> 1) * D.ncol, cp, scale, q);
}

// inverse FFT of one plane (decimation in time, bit-reversed order in, natural order out) -> phi of levels 2p (real part) and 2p + 1
template <typename Real, int LX2 = 0, int LY2 = 0>
__global__ void __launch_bounds__(LX2 ? FFT_NT : 1024)
g3_ifft_kernel(Dims D_arg, const cx<Real>* Z_all, Real* phi_all, const cx<Real>* twx, const cx<Real>* twy, const int* env_ids)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<Real>* Z = reinterpret_cast<cx<Real>*>(smem_raw);
    const Dims D = plane_dims<LX2, LY2>(D_arg);
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y, pz = blockIdx.x, nzp = (D.nz + 1) >> 1;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const Real* in = reinterpret_cast<const Real*>(Z_all + ((size_t)env * nzp + pz) * D.ncol);
    const int pitch = D.nx + 1;      // odd row pitch: see plane_fft_forward
#pragma unroll 2
    for (int c = 2 * threadIdx.x; c < D.ncol; c += 2 * blockDim.x) {
        const Quad<Real> q = ld4(in + 2 * c);
        const int zi = (c >> D.lx2) * pitch + (c & (D.nx - 1));
        Z[zi] = cx<Real>{q.v[0], q.v[1]}; Z[zi + 1] = cx<Real>{q.v[2], q.v[3]};
    }
    cx<Real>* tw = Z + D.ny * pitch;
    stage_twiddles<Real>(D, tw, twx, twy);
    __syncthreads();
    if (LX2) plane_fft_inverse<Real>(D, Z, tw, tw + (D.nx >> 1), BlockRunFixed<FFT_NT>{}, pitch, true);
    else plane_fft_inverse<Real>(D, Z, tw, tw + (D.nx >> 1), BlockRun{}, pitch, true);
    const Real norm = Real(1) / (Real)D.ncol;
    Real* phi = phi_all + ((size_t)env * D.nz + 2 * pz) * D.ncol;
    const bool odd = 2 * pz + 1 < D.nz;
    for (int c = 4 * threadIdx.x; c < D.ncol; c += 4 * blockDim.x) {
        const int zi = (c >> D.lx2) * pitch + (c & (D.nx - 1));
        const cx<Real> z0 = Z[zi], z1 = Z[zi + 1], z2 = Z[zi + 2], z3 = Z[zi + 3];
        st4(phi + c, Quad<Real>{{z0.re * norm, z1.re * norm, z2.re * norm, z3.re * norm}});
        if (odd) st4(phi + D.ncol + c, Quad<Real>{{z0.im * norm, z1.im * norm, z2.im * norm, z3.im * norm}});
    }
}

template <typename Real>
__global__ void __launch_bounds__(TB)
g3_correct_kernel(Dims D, ConstsG<Real> C, Real* P_all, const Real* phi_all, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= D.nc) return;
    cell_correct<Real>(D, C, P_all + (size_t)env * D.nstate, phi_all + (size_t)env * D.nc, cell);
}
// the same for four consecutive cells of a row per thread (the expressions of cell_correct; vector loads and stores)
template <typename Real>
__global__ void __launch_bounds__(TB)
g3_correct4_kernel(Dims D, ConstsG<Real> C, Real* P_all, const Real* phi_all, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (env < 0) return;      // an environment the fused vector step only re-initialises (next_step mode)
    const int cell = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (cell >= D.nc) return;
    Real* P = P_all + (size_t)env * D.nstate;
    const Real* phi = phi_all + (size_t)env * D.nc;
    const int i = cell & (D.nx - 1), j = (cell >> D.lx2) & (D.ny - 1), k = cell >> (D.lx2 + D.ly2);
    const Quad<Real> ph = ld4(phi + cell), pj = ld4(phi + (k * D.ny + ((j - 1) & (D.ny - 1))) * D.nx + i);
    const Real pl = phi[cell - i + ((i - 1) & (D.nx - 1))];
    Quad<Real> u = ld4(P + D.gu + cell), v = ld4(P + D.gv + cell);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        u.v[m] -= (ph.v[m] - (m ? ph.v[m - 1] : pl)) * C.idx;
        v.v[m] -= (ph.v[m] - pj.v[m]) * C.idy;
    }
    st4(P + D.gu + cell, u);
    st4(P + D.gv + cell, v);
    if (k >= 1) {
        const Quad<Real> pk = ld4(phi + cell - D.ncol);
        Quad<Real> w = ld4(P + D.gw + cell);
#pragma unroll
        for (int m = 0; m < 4; ++m) w.v[m] -= (ph.v[m] - pk.v[m]) * C.idz;
        st4(P + D.gw + cell, w);
    }
}

// epilogue, part 1: Nusselt sum and NaN count of a chunk of cells (one partial per block into acc[env][block][2], added up in block
// order by part 2: the Nusselt number — the reward — is reproducible bit for bit, which atomics would not give); write-back to the environment's
// own array when the march ended in the other buffer; observation = get_state (rbc_sim3D_api.jl:106-121, rbc3D.py:229-232)
template <typename Real>
__global__ void __launch_bounds__(TB)
g3_reduce_kernel(Dims D, ConstsG<Real> C, const Real* cur_all, Real* st_all, float* obs, double* acc, const int* env_ids)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    const Real* cur = cur_all + (size_t)env * D.nstate;
    Real* st = st_all + (size_t)env * D.nstate;
    double a = 0, bad = 0;
    for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < D.nc; cell += gridDim.x * blockDim.x) {
        const int k = cell / D.ncol;
        const double b = (double)cur[D.gb + cell], u = (double)cur[D.gu + cell], v = (double)cur[D.gv + cell], w = (double)cur[D.gw + cell];
        if (b != b || u != u || v != v || w != w) bad += 1;
        // get_nusselt (rbc_sim3D_api.jl:134-159): conduction profile on a UNIT height, z = (k + 1/2)/nz
        const double zc = (k + 0.5) / D.nz, Tc = (1.0 - zc) * C.delta_b_d + C.b_top_d;
        a += (b - Tc) * w;
    }
    if (cur != st)
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < D.nstate; q += gridDim.x * blockDim.x) st[q] = cur[q];
    if (obs != nullptr) {
        float* ob = obs + (size_t)env * 4 * D.nc;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 4 * D.nc; q += gridDim.x * blockDim.x) ob[q] = (float)cur[q];
    }
    __shared__ double sa[TB], sb[TB];
    sa[threadIdx.x] = a; sb[threadIdx.x] = bad;
    __syncthreads();
    for (int s = TB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { acc[((size_t)env * gridDim.x + blockIdx.x) * 2] = sa[0]; acc[((size_t)env * gridDim.x + blockIdx.x) * 2 + 1] = sb[0]; }
}

// epilogue, part 2: Nusselt number, reward, NaN flag, clock and truncation of every listed environment
template <typename Real>
__global__ void g3_finalize_kernel(Dims D, ConstsG<Real> C, const double* kappa_env, const double* acc, int nblocks, rbc3dg_api::IoRaw io,
                                   const int* env_ids, int n, int advance_clock)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int env = env_ids ? env_ids[j] : j;
    double sum = 0, bad = 0;
    for (int b = 0; b < nblocks; ++b) { sum += acc[((size_t)env * nblocks + b) * 2]; bad += acc[((size_t)env * nblocks + b) * 2 + 1]; }
    const double nu = 1.0 + (sum / (double)D.nc) / kappa_env[env];
    io.nusselt[env] = nu;
    io.reward[env] = (float)(-nu);
    io.nan_flag[env] = bad > 0 ? 1 : 0;
    if (advance_clock) {
        const double tn = io.t[env] + C.dt_action;
        io.t[env] = tn;
        io.step_count[env] += 1;
        io.truncated[env] = tn >= C.episode_length ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// The fused vector step on this path (rbc2d::VecIO; the semantics of env_epilogue3 / env_action_step3 in rbc3d_core.h, statement for
// statement, spread over batch kernels): next_step environments that truncated on the previous call are re-initialised from the
// checkpoint bank and skip the march; same_step truncations and NaN resets store the terminal observation / Nusselt number / return,
// gather a new state and run the observation epilogue a second time.
// ------------------------------------------------------------------------------------------
__global__ void g3_vec_begin_kernel(const int* env_ids, int n, const int* pending, int* list)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int env = env_ids ? env_ids[j] : j;
    list[j] = pending[env] ? -1 : env;
}
// state <- bank[checkpoint_draw(seed, id_offset + env, episode[env])] for the environments with which[env] != 0
template <typename Real>
__global__ void __launch_bounds__(TB)
g3_vec_gather_kernel(Dims D, Real* st_all, rbc2d::VecIO V, const int* env_ids, const int* which)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (!which[env]) return;
    const double* src = V.bank + (size_t)rbc2d::checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * D.nstate;
    Real* st = st_all + (size_t)env * D.nstate;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < D.nstate; q += gridDim.x * blockDim.x) st[q] = (Real)src[q];
}
// Nusselt / NaN partials like g3_reduce_kernel, from the environment's own array where the march was skipped (`from_state`) and only
// for the environments marked in `only` (second pass); no observation
template <typename Real>
__global__ void __launch_bounds__(TB)
g3_vec_reduce_kernel(Dims D, ConstsG<Real> C, const Real* cur_all, Real* st_all, double* acc, const int* env_ids, const int* from_state, const int* only)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (only && !only[env]) return;
    Real* st = st_all + (size_t)env * D.nstate;
    const Real* cur = ((from_state && from_state[env]) || only) ? st : cur_all + (size_t)env * D.nstate;
    double a = 0, bad = 0;
    for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < D.nc; cell += gridDim.x * blockDim.x) {
        const int k = cell / D.ncol;
        const double b = (double)cur[D.gb + cell], u = (double)cur[D.gu + cell], v = (double)cur[D.gv + cell], w = (double)cur[D.gw + cell];
        if (b != b || u != u || v != v || w != w) bad += 1;
        const double zc = (k + 0.5) / D.nz, Tc = (1.0 - zc) * C.delta_b_d + C.b_top_d;
        a += (b - Tc) * w;
    }
    if (cur != st)
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < D.nstate; q += gridDim.x * blockDim.x) st[q] = cur[q];
    __shared__ double sa[TB], sb[TB];
    sa[threadIdx.x] = a; sb[threadIdx.x] = bad;
    __syncthreads();
    for (int s = TB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { acc[((size_t)env * gridDim.x + blockIdx.x) * 2] = sa[0]; acc[((size_t)env * gridDim.x + blockIdx.x) * 2 + 1] = sb[0]; }
}
// the bookkeeping of env_epilogue3 (first pass): flags, reward, return, clock, pending marks; do_reset[env] for the kernels behind it
template <typename Real>
__global__ void g3_vec_finalize_kernel(Dims D, ConstsG<Real> C, const double* kappa_env, const double* acc, int nblocks, rbc3dg_api::IoRaw io,
                                       rbc2d::VecIO V, const int* env_ids, int n, int* do_reset_out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int env = env_ids ? env_ids[j] : j;
    double sum = 0, nbad = 0;
    for (int b = 0; b < nblocks; ++b) { sum += acc[((size_t)env * nblocks + b) * 2]; nbad += acc[((size_t)env * nblocks + b) * 2 + 1]; }
    const double nu = 1.0 + (sum / (double)D.nc) / kappa_env[env];
    const bool bad = nbad > 0;
    const int pend = (V.mode == 1 && V.bank != nullptr) ? V.pending[env] : 0;
    const double t_old = io.t[env];
    const bool stepping = !pend;
    const bool trunc_now = stepping && (t_old + C.dt_action >= C.episode_length);
    const bool bad_reset = stepping && V.nan_reset && bad;
    const bool do_reset = stepping && V.bank != nullptr && ((V.mode == 2 && trunc_now) || bad_reset);
    do_reset_out[env] = do_reset ? 1 : 0;
    if (pend) {
        io.nusselt[env] = nu;
        io.reward[env] = 0.0f;
        io.nan_flag[env] = bad ? 1 : 0;
        io.truncated[env] = 0;
        io.t[env] = 0.0;
        io.step_count[env] = 1;
        V.episode[env] += 1;
        V.ep_return[env] = 0.0;
        V.pending[env] = 0;
        return;
    }
    double rew = -nu;
    if (bad_reset) rew = 0.0;
    if (bad && V.nan_count != nullptr) atomicAdd(V.nan_count, 1);
    io.reward[env] = (float)rew;
    io.nan_flag[env] = bad ? 1 : 0;
    const double ret = V.ep_return[env] + (double)(float)rew;
    if (do_reset) {
        if (V.final_nu_a != nullptr) V.final_nu_a[env] = nu;
        if (V.final_return != nullptr) V.final_return[env] = ret;
        io.truncated[env] = 1;
    } else {
        io.nusselt[env] = nu;
        io.t[env] = t_old + C.dt_action;
        io.step_count[env] += 1;
        io.truncated[env] = (trunc_now || bad_reset) ? 1 : 0;
        V.ep_return[env] = ret;
        if (V.pending != nullptr) V.pending[env] = (V.mode == 1 && trunc_now && V.bank != nullptr) ? 1 : 0;
    }
}
// observation (= get_state) of every listed environment; an environment that is re-initialised inside this step stores it as its
// terminal observation instead and takes its new state from the bank (element by element: read, emit, overwrite)
template <typename Real>
__global__ void __launch_bounds__(TB)
g3_vec_emit_kernel(Dims D, Real* st_all, float* obs, rbc2d::VecIO V, const int* env_ids, const int* do_reset)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    Real* st = st_all + (size_t)env * D.nstate;
    if (!do_reset[env]) {
        if (obs == nullptr) return;
        float* ob = obs + (size_t)env * 4 * D.nc;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 4 * D.nc; q += gridDim.x * blockDim.x) ob[q] = (float)st[q];
        return;
    }
    const double* src = V.bank + (size_t)rbc2d::checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * D.nstate;
    float* fo = V.final_obs ? V.final_obs + (size_t)env * 4 * D.nc : nullptr;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < D.nstate; q += gridDim.x * blockDim.x) {
        if (fo != nullptr && q < 4 * D.nc) fo[q] = (float)st[q];
        st[q] = (Real)src[q];
    }
}
// second pass of the environments re-initialised inside this step: Nusselt number and observation of the new state, fresh episode
template <typename Real>
__global__ void g3_vec_finalize2_kernel(Dims D, const double* kappa_env, const double* acc, int nblocks, rbc3dg_api::IoRaw io, rbc2d::VecIO V,
                                        const int* env_ids, int n, const int* do_reset)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int env = env_ids ? env_ids[j] : j;
    if (!do_reset[env]) return;
    double sum = 0;
    for (int b = 0; b < nblocks; ++b) sum += acc[((size_t)env * nblocks + b) * 2];
    io.nusselt[env] = 1.0 + (sum / (double)D.nc) / kappa_env[env];
    io.t[env] = 0.0;
    io.step_count[env] = 1;
    V.episode[env] += 1;
    V.ep_return[env] = 0.0;
    if (V.pending != nullptr) V.pending[env] = 0;
}
template <typename Real>
__global__ void __launch_bounds__(TB)
g3_vec_emit2_kernel(Dims D, const Real* st_all, float* obs, const int* env_ids, const int* do_reset)
{
    const int env = env_ids ? env_ids[blockIdx.y] : blockIdx.y;
    if (!do_reset[env] || obs == nullptr) return;
    const Real* st = st_all + (size_t)env * D.nstate;
    float* ob = obs + (size_t)env * 4 * D.nc;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 4 * D.nc; q += gridDim.x * blockDim.x) ob[q] = (float)st[q];
}

}  // namespace

namespace rbc3dg_api {

constexpr int MAX_CHAINS = 8;
constexpr int MAX_RBLOCKS = 128;      // blocks per environment of the epilogue reduction
struct Plan {
    Dims D;
    HostConfigG hc;
    int B = 0, precision = 32, device = 0;
    size_t rs = 4, smem = 0;
    int fixed_plane = 0;         // 66 / 55: the plane kernels run their instantiation with compile-time extents (64 x 64 / 32 x 32 columns); RBC_B200_G3_FIXED_PLANE=0: run-time extents
    int fft_threads = 256;       // with the stages fused two at a time: 256 threads 860, 512 threads 854, 1024 threads 813 env-steps/s (RBC_B200_G3_FFT_THREADS)
    int variant = 0;             // RBC_B200_G3_VARIANT: wall-order variants of the parity study (rbc3dg_core.h); 0 = the scheme of record
    int tiled = 0;               // rows of the column patch of the tiled tendency (shared-memory plane ring; nx % 32 == 0, ny % rows == 0); 0 = per-cell kernel.  RBC_B200_G3_TILED=0|8|16 overrides
    void *P = nullptr, *G = nullptr, *Z = nullptr, *phi = nullptr, *Tb = nullptr, *cp = nullptr, *twx = nullptr, *twy = nullptr;
    void *nu = nullptr, *kappa = nullptr;          // per-environment diffusivities (Real)
    double* kappa_d = nullptr;                      // the same in fp64 for the Nusselt number
    double* acc = nullptr;
    // Independent chains on several streams.  Every kernel of a stage covers the whole batch with a grid that is rarely a multiple of
    // what the GPU holds at once (64 environments at 64 x 64 x 32: 1024 FFT planes on 888 resident CTAs, 1024 tendency patches on
    // 296), so the last wave of each launch leaves SMs idle until the next kernel of the chain may start.  With the batch cut into
    // independent chains the block scheduler fills the tail of one chain's kernel with another chain's next one.  Environments never
    // interact, so results do not depend on the cut.  Measured at 64 x 64 x 32 (env-steps/s with 1 / 2 / 3 / 4 chains, build of that day):
    // 64 environments 1125 / 1187 / 1208 / 1241; 148 environments 1224 / 1267 / - / 1266; 14 environments 772 / 792 / - / 778.  Final build,
    // 4 / 6 / 8 chains: 64 environments 1371 / 1432 / 1436, 148 environments 1413 / - / 1397.
    int streams = 0;             // 0 = 8 chains for 16 .. 96 environments, 4 above, 2 below; RBC_B200_G3_STREAMS=1..8 fixes the number
    int* iota = nullptr;         // 0 .. B-1: the environment list of a launch over the whole batch, so that it can be cut
    int* vlist = nullptr;        // fused vector step: the march's environment list with -1 for environments that only re-initialise
    int* do_reset = nullptr;     // fused vector step: environments re-initialised inside the current step
    cudaStream_t side[MAX_CHAINS - 1] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHAINS - 1] = {};
};

int supported(int nx, int ny, int nz) { return dims_supported(nx, ny, nz) ? 1 : 0; }
size_t smem_bytes(const Plan* p) { return p->smem; }
int values_per_env(const Plan* p) { return p->D.nstate; }

template <typename Real>
static int upload_tables(Plan* p)
{
    const Dims& D = p->D;
    std::vector<double> cp((size_t)D.nz * D.ncol), tx(D.nx), ty(D.ny);
    build_pivots_host(D, p->hc.lx, p->hc.ly, p->hc.lz, cp.data());
    build_twiddles_host(D.nx, tx.data());
    build_twiddles_host(D.ny, ty.data());
    std::vector<Real> a(cp.begin(), cp.end()), b(tx.begin(), tx.end()), c(ty.begin(), ty.end());
    CK(cudaMalloc(&p->cp, a.size() * sizeof(Real)));
    CK(cudaMalloc(&p->twx, b.size() * sizeof(Real)));
    CK(cudaMalloc(&p->twy, c.size() * sizeof(Real)));
    CK(cudaMemcpy(p->cp, a.data(), a.size() * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->twx, b.data(), b.size() * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->twy, c.data(), c.size() * sizeof(Real), cudaMemcpyHostToDevice));
    return 0;
}

template <typename Real>
static int upload_rayleigh(Plan* p, const double* ra /*[B] or nullptr*/)
{
    std::vector<Real> nu(p->B), ka(p->B);
    std::vector<double> kd(p->B);
    for (int e = 0; e < p->B; ++e) {
        const double r = ra ? ra[e] : p->hc.ra;
        if (!(r > 0)) return rbc_fail("rbc3d: Rayleigh numbers must be positive");
        nu[e] = (Real)sqrt(p->hc.pr / r); kd[e] = 1.0 / sqrt(p->hc.pr * r); ka[e] = (Real)kd[e];       // rbc_sim3D_api.jl:37-38
    }
    CK(cudaMemcpy(p->nu, nu.data(), p->B * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->kappa, ka.data(), p->B * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->kappa_d, kd.data(), p->B * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

int set_rayleigh(Plan* p, const double* ra_host)
{
    CK(cudaSetDevice(p->device));
    return p->precision == 32 ? upload_rayleigh<float>(p, ra_host) : upload_rayleigh<double>(p, ra_host);
}

int create(const HostConfigG& hc, int nx, int ny, int nz, int num_envs, int precision, int device, Plan** out)
{
    *out = nullptr;
    if (!dims_supported(nx, ny, nz)) return rbc_fail("rbc3d_create: grid must have nx, ny powers of two in 8..256 and 6 <= nz <= 256");
    Plan* p = new Plan();
    p->D = make_dims(nx, ny, nz);
    p->hc = hc; p->B = num_envs; p->precision = precision; p->device = device;
    p->rs = precision == 32 ? 4 : 8;
    p->smem = ((size_t)p->D.ny * (p->D.nx + 1) + p->D.nx / 2 + p->D.ny / 2) * 2 * p->rs;      // one complex plane (odd row pitch) + the twiddles of both directions
    if (p->smem > 227 * 1024) { delete p; return rbc_fail("rbc3d_create: a horizontal plane of this grid does not fit the shared memory of an SM"); }
    cudaError_t e = cudaSuccess;
    {
        const char* fx = getenv("RBC_B200_G3_FIXED_PLANE");
        if (!(fx && atoi(fx) == 0)) p->fixed_plane = (nx == 64 && ny == 64) ? 66 : ((nx == 32 && ny == 32) ? 55 : 0);
        const void* fns[6] = {precision == 32 ? (const void*)g3_div_fft_kernel<float> : (const void*)g3_div_fft_kernel<double>,
                              precision == 32 ? (const void*)g3_ifft_kernel<float> : (const void*)g3_ifft_kernel<double>,
                              precision == 32 ? (const void*)g3_div_fft_kernel<float, 6, 6> : (const void*)g3_div_fft_kernel<double, 6, 6>,
                              precision == 32 ? (const void*)g3_ifft_kernel<float, 6, 6> : (const void*)g3_ifft_kernel<double, 6, 6>,
                              precision == 32 ? (const void*)g3_div_fft_kernel<float, 5, 5> : (const void*)g3_div_fft_kernel<double, 5, 5>,
                              precision == 32 ? (const void*)g3_ifft_kernel<float, 5, 5> : (const void*)g3_ifft_kernel<double, 5, 5>};
        for (int q = 0; q < 6 && e == cudaSuccess; ++q) e = cudaFuncSetAttribute(fns[q], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
    }
    if (e != cudaSuccess) { delete p; return rbc_fail(std::string("rbc3d_create: ") + cudaGetErrorString(e)); }
    {
        if (const char* v = getenv("RBC_B200_G3_VARIANT")) p->variant = atoi(v);
        if (const char* v = getenv("RBC_B200_G3_FFT_THREADS")) { const int t = atoi(v); if (t >= 64 && t <= 1024 && t % 32 == 0) p->fft_threads = t; }
        const char* sw = getenv("RBC_B200_G3_TILED");
        const int want = sw ? atoi(sw) : 8;       // 8 rows: two CTAs per SM hide each other's barriers and copy waits (678 vs 634 env-steps/s at 64 x 64 x 32)
        p->tiled = (nx % TT_X != 0 || want == 0) ? 0 : ((want >= 16 && ny % 16 == 0) ? 16 : (ny % 8 == 0 ? 8 : 0));
        if (p->tiled) {
            const int bytes = (int)((size_t)PIPE_R * (p->tiled == 16 ? TileGeom<16>::SLOT : TileGeom<8>::SLOT) * p->rs + 16 * PIPE_R);      // ring + mbarriers
            const void* fn = precision == 32 ? (p->tiled == 16 ? (const void*)g3_tendency_tiled_kernel<float, 16> : (const void*)g3_tendency_tiled_kernel<float, 8>)
                                             : (p->tiled == 16 ? (const void*)g3_tendency_tiled_kernel<double, 16> : (const void*)g3_tendency_tiled_kernel<double, 8>);
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            if (e != cudaSuccess) { delete p; return rbc_fail(std::string("rbc3d_create: ") + cudaGetErrorString(e)); }
        }
    }
    const size_t B = num_envs, rs = p->rs;
    const Dims& D = p->D;
    struct { void** ptr; size_t bytes; } allocs[] = {
        {&p->P, B * D.nstate * rs}, {&p->G, B * 4 * D.nc * rs}, {&p->Z, B * D.nc * 2 * rs}, {&p->phi, B * D.nc * rs},
        {&p->Tb, B * D.ncol * rs}, {&p->nu, B * rs}, {&p->kappa, B * rs}, {(void**)&p->kappa_d, B * sizeof(double)},
        {(void**)&p->acc, B * MAX_RBLOCKS * 2 * sizeof(double)}};
    for (auto& a : allocs) {
        e = cudaMalloc(a.ptr, a.bytes);
        if (e != cudaSuccess) { destroy(p); return rbc_fail(std::string("rbc3d_create: cudaMalloc: ") + cudaGetErrorString(e)); }
        cudaMemset(*a.ptr, 0, a.bytes);
    }
    {
        if (const char* v = getenv("RBC_B200_G3_STREAMS")) { const int t = atoi(v); p->streams = t < 0 ? 0 : (t > MAX_CHAINS ? MAX_CHAINS : t); }
        std::vector<int> io(num_envs);
        for (int q = 0; q < num_envs; ++q) io[q] = q;
        e = cudaMalloc((void**)&p->iota, B * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc((void**)&p->vlist, B * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc((void**)&p->do_reset, B * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(p->do_reset, 0, B * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(p->iota, io.data(), B * sizeof(int), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
        for (int c = 0; c < MAX_CHAINS - 1 && e == cudaSuccess; ++c) {
            e = cudaStreamCreateWithFlags(&p->side[c], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_join[c], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) { destroy(p); return rbc_fail(std::string("rbc3d_create: ") + cudaGetErrorString(e)); }
    }
    int rc = precision == 32 ? upload_tables<float>(p) : upload_tables<double>(p);
    if (!rc) rc = set_rayleigh(p, nullptr);
    if (rc) { destroy(p); return rc; }
    *out = p;
    return 0;
}

void destroy(Plan* p)
{
    if (!p) return;
    void* ptrs[] = {p->P, p->G, p->Z, p->phi, p->Tb, p->cp, p->twx, p->twy, p->nu, p->kappa, p->kappa_d, p->acc};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (p->iota) cudaFree(p->iota);
    if (p->vlist) cudaFree(p->vlist);
    if (p->do_reset) cudaFree(p->do_reset);
    for (int c = 0; c < MAX_CHAINS - 1; ++c) {
        if (p->side[c]) cudaStreamDestroy(p->side[c]);
        if (p->ev_join[c]) cudaEventDestroy(p->ev_join[c]);
    }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    delete p;
}

template <typename Real>
static int project_t(Plan* p, const ConstsG<Real>& C, Real* buf, const int* env_ids, int n, cudaStream_t st, int64_t* launches)
{
    const Dims& D = p->D;
    const dim3 gplane((D.nz + 1) / 2, n), gcell((D.nc + TB - 1) / TB, n), gmode((D.ncol + 127) / 128, n);
    const Real dz = (Real)(p->hc.lz / D.nz);
    cx<Real>* Z = (cx<Real>*)p->Z;
    const cx<Real>*twx = (const cx<Real>*)p->twx, *twy = (const cx<Real>*)p->twy;
    if (p->fixed_plane == 66) g3_div_fft_kernel<Real, 6, 6><<<gplane, FFT_NT, p->smem, st>>>(D, C, buf, Z, twx, twy, env_ids);
    else if (p->fixed_plane == 55) g3_div_fft_kernel<Real, 5, 5><<<gplane, FFT_NT, p->smem, st>>>(D, C, buf, Z, twx, twy, env_ids);
    else g3_div_fft_kernel<Real><<<gplane, p->fft_threads, p->smem, st>>>(D, C, buf, Z, twx, twy, env_ids);
    g3_thomas_kernel<Real><<<gmode, 128, 0, st>>>(D, Z, (const Real*)p->cp, dz * dz, env_ids);
    if (p->fixed_plane == 66) g3_ifft_kernel<Real, 6, 6><<<gplane, FFT_NT, p->smem, st>>>(D, Z, (Real*)p->phi, twx, twy, env_ids);
    else if (p->fixed_plane == 55) g3_ifft_kernel<Real, 5, 5><<<gplane, FFT_NT, p->smem, st>>>(D, Z, (Real*)p->phi, twx, twy, env_ids);
    else g3_ifft_kernel<Real><<<gplane, p->fft_threads, p->smem, st>>>(D, Z, (Real*)p->phi, twx, twy, env_ids);
    if ((reinterpret_cast<size_t>(buf) & 15) == 0)
        g3_correct4_kernel<Real><<<dim3((D.nc / 4 + TB - 1) / TB, n), TB, 0, st>>>(D, C, buf, (const Real*)p->phi, env_ids);
    else
        g3_correct_kernel<Real><<<gcell, TB, 0, st>>>(D, C, buf, (const Real*)p->phi, env_ids);
    CK(cudaGetLastError());
    *launches += 4;
    return 0;
}

// the tendency + RK3 substep of one stage for the listed environments
template <typename Real>
static void tendency_t(Plan* p, const ConstsG<Real>& C, const Real* cur, Real* nxt, const int* env_ids, int n, Real dt, Real gam, Real zet, int stage,
                       cudaStream_t st, int64_t* launches)
{
    const Dims& D = p->D;
    const size_t sm16 = (size_t)PIPE_R * TileGeom<16>::SLOT * sizeof(Real) + 16 * PIPE_R, sm8 = (size_t)PIPE_R * TileGeom<8>::SLOT * sizeof(Real) + 16 * PIPE_R;
    if (p->tiled == 16)
        g3_tendency_tiled_kernel<Real, 16><<<dim3((D.nx / TT_X) * (D.ny / 16), n), TileGeom<16>::NT, sm16, st>>>(
            D, C, (const Real*)p->nu, (const Real*)p->kappa, cur, nxt, (Real*)p->G, (const Real*)p->Tb, env_ids, dt, gam, zet, stage > 0, stage < 2);
    else if (p->tiled == 8)
        g3_tendency_tiled_kernel<Real, 8><<<dim3((D.nx / TT_X) * (D.ny / 8), n), TileGeom<8>::NT, sm8, st>>>(
            D, C, (const Real*)p->nu, (const Real*)p->kappa, cur, nxt, (Real*)p->G, (const Real*)p->Tb, env_ids, dt, gam, zet, stage > 0, stage < 2);
    else
        g3_tendency_kernel<Real><<<dim3((D.nc + TB - 1) / TB, n), TB, 0, st>>>(D, C, (const Real*)p->nu, (const Real*)p->kappa, cur, nxt, (Real*)p->G,
                                                                                 (const Real*)p->Tb, env_ids, dt, gam, zet, stage > 0, stage < 2);
    *launches += 1;
}

template <typename Real>
static int launch_t(Plan* p, const IoRaw& io, const int* env_ids, int n, int nsub, int project_first, int advance_clock, cudaStream_t st,
                    int64_t* launches, const rbc2d::VecIO* vecp)
{
    const Dims& D = p->D;
    ConstsG<Real> C = make_consts<Real>(D, p->hc);
    C.variant = p->variant;
    const Real gam[3] = {Real(8.0 / 15.0), Real(5.0 / 12.0), Real(3.0 / 4.0)};
    const Real zet[3] = {Real(0), Real(-17.0 / 60.0), Real(-5.0 / 12.0)};
    Real* S = (Real*)io.state;
    Real* cur = S;
    Real* nxt = (Real*)p->P;
    // fused vector step (rbc2d::VecIO): next_step environments that truncated on the previous call take a new state from the bank and
    // are struck from the march's list
    const bool vec = vecp != nullptr && vecp->mode >= 0 && advance_clock;
    const rbc2d::VecIO V = vec ? *vecp : rbc2d::VecIO();
    const bool skip_pending = vec && V.mode == 1 && V.bank != nullptr;
    const int eblocks = (D.nstate + 4 * TB - 1) / (4 * TB) < 128 ? (D.nstate + 4 * TB - 1) / (4 * TB) : 128;
    const int* const listed = env_ids;          // the environments of this call (the epilogue covers all of them)
    if (skip_pending) {
        g3_vec_begin_kernel<<<(n + 127) / 128, 128, 0, st>>>(listed, n, V.pending, p->vlist);
        g3_vec_gather_kernel<Real><<<dim3(eblocks, n), TB, 0, st>>>(D, S, V, listed, V.pending);
        *launches += 2;
        env_ids = p->vlist;
    }
    // the chains: the whole list on the caller's stream, or equal parts of it on the caller's stream and the plan's side streams
    const int want = p->streams > 0 ? p->streams : (n > 96 ? 4 : (n >= 16 ? 8 : 2));
    const int nch = want < n ? want : (n < 1 ? 1 : n);
    const int* ids[MAX_CHAINS] = {env_ids};
    int cnt[MAX_CHAINS] = {n};
    cudaStream_t strm[MAX_CHAINS] = {st};
    if (nch > 1) {
        const int* all = env_ids ? env_ids : p->iota;
        CK(cudaEventRecord(p->ev_fork, st));
        for (int c = 0, at = 0; c < nch; ++c) {
            cnt[c] = n / nch + (c < n % nch ? 1 : 0);
            ids[c] = all + at; at += cnt[c];
            if (c > 0) { strm[c] = p->side[c - 1]; CK(cudaStreamWaitEvent(strm[c], p->ev_fork, 0)); }
        }
    }
    if (nsub > 0)
        for (int c = 0; c < nch; ++c) {
            g3_heater_kernel<Real><<<dim3((D.ncol + TB - 1) / TB, cnt[c]), TB, 0, strm[c]>>>(D, C, io.actions, (Real*)p->Tb, ids[c]);
            *launches += 1;
        }
    if (project_first)
        for (int c = 0; c < nch; ++c) {
            int rc = project_t<Real>(p, C, cur, ids[c], cnt[c], strm[c], launches);
            if (rc) return rc;
        }
    for (int sub = 0; sub < nsub; ++sub) {
        const Real dt = (sub == nsub - 1) ? C.dt_last : C.dt_full;
        for (int stage = 0; stage < 3; ++stage) {
            for (int c = 0; c < nch; ++c) {
                tendency_t<Real>(p, C, cur, nxt, ids[c], cnt[c], dt, gam[stage], zet[stage], stage, strm[c], launches);
                int rc = project_t<Real>(p, C, nxt, ids[c], cnt[c], strm[c], launches);
                if (rc) return rc;
            }
            Real* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
    for (int c = 1; c < nch; ++c) {
        CK(cudaEventRecord(p->ev_join[c - 1], strm[c]));
        CK(cudaStreamWaitEvent(st, p->ev_join[c - 1], 0));
    }
    const int rblocks = (D.nc + 4 * TB - 1) / (4 * TB) < MAX_RBLOCKS ? (D.nc + 4 * TB - 1) / (4 * TB) : MAX_RBLOCKS;
    if (!vec) {
        g3_reduce_kernel<Real><<<dim3(rblocks, n), TB, 0, st>>>(D, C, cur, S, io.obs, p->acc, env_ids);
        g3_finalize_kernel<Real><<<(n + 127) / 128, 128, 0, st>>>(D, C, p->kappa_d, p->acc, rblocks, io, env_ids, n, advance_clock);
        CK(cudaGetLastError());
        *launches += 2;
        return 0;
    }
    g3_vec_reduce_kernel<Real><<<dim3(rblocks, n), TB, 0, st>>>(D, C, cur, S, p->acc, listed, skip_pending ? V.pending : nullptr, nullptr);
    g3_vec_finalize_kernel<Real><<<(n + 127) / 128, 128, 0, st>>>(D, C, p->kappa_d, p->acc, rblocks, io, V, listed, n, p->do_reset);
    g3_vec_emit_kernel<Real><<<dim3(eblocks, n), TB, 0, st>>>(D, S, io.obs, V, listed, p->do_reset);
    *launches += 3;
    if (V.bank != nullptr && (V.mode == 2 || V.nan_reset)) {       // second pass of the environments re-initialised inside this step
        g3_vec_reduce_kernel<Real><<<dim3(rblocks, n), TB, 0, st>>>(D, C, S, S, p->acc, listed, nullptr, p->do_reset);
        g3_vec_finalize2_kernel<Real><<<(n + 127) / 128, 128, 0, st>>>(D, p->kappa_d, p->acc, rblocks, io, V, listed, n, p->do_reset);
        g3_vec_emit2_kernel<Real><<<dim3(eblocks, n), TB, 0, st>>>(D, S, io.obs, listed, p->do_reset);
        *launches += 3;
    }
    CK(cudaGetLastError());
    return 0;
}

int launch(Plan* p, const IoRaw& io, const int* env_ids, int n, int nsub, int project_first, int advance_clock, cudaStream_t stream,
           int64_t* launches, const rbc2d::VecIO* vec)
{
    if (n <= 0) return 0;
    if (n > 65535) return rbc_fail("rbc3d: at most 65535 environments per launch on this path");
    return p->precision == 32 ? launch_t<float>(p, io, env_ids, n, nsub, project_first, advance_clock, stream, launches, vec)
                              : launch_t<double>(p, io, env_ids, n, nsub, project_first, advance_clock, stream, launches, vec);
}

int nsub_of(const Plan* p)
{
    return make_consts<float>(p->D, p->hc).nsub;
}

}  // namespace rbc3dg_api
