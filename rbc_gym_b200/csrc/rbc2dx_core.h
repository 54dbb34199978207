// rbc2dx_core.h — the 2D Rayleigh-Benard action step on ANY registered grid, one thread-block CLUSTER
// per environment.
//
// Same reference path as rbc2d_core.h (step_simulation -> run!, src/rbc_gym/sim/rbc_sim2D_api.jl:75-97;
// model configuration src/rbc_gym/sim/rbc_sim2D.jl:149-160; scheme per SURVEY.md 8a), generalised from the
// registered 96 x 64 grid to the grids RBC-Gym is run on at higher Rayleigh numbers (config 3: 192 x 128).
//
// A 192 x 128 environment (296 KB in fp32) does not fit the shared memory of one SM, so the environment is
// split in z over the CL CTAs of a cluster; every CTA keeps its slab of NZL = NZ/CL rows (plus 3 halo rows
// on each side) on-chip for all RK3 stages of an action step.  What crosses CTAs goes through distributed
// shared memory:
//   * stencil halos    — pushed into the neighbours' halo rows by the thread that produces the value
//                        (b* and the shared w* face during the tendency march, u and w after the pressure
//                        correction) with st.async: fire-and-forget remote stores that complete a transaction
//                        count on an mbarrier of the RECEIVING CTA, which waits on its own barrier just before it
//                        consumes the rows — no cluster-wide barrier and no memory fence in the stage loop;
//   * the z-tridiagonal Poisson solve — SPIKE partitioning: every CTA solves its own NZL-row block with the
//                        two-sided Thomas sweep, the 2*CL block-end values are exchanged, each CTA applies the
//                        pre-inverted reduced system (host, fp64) and corrects its block with the two
//                        precomputed spike vectors.  The reduced solve also yields the pressure row just
//                        below the slab, which is inverse-transformed locally (idle FFT lanes), so the
//                        pressure correction needs no further exchange;
//   * the epilogue reductions (Nusselt numbers, NaN flag).
// The x-direction (FFT, upwind stencils) never leaves the CTA.  In fp32 mode the stage loop has no cluster
// barrier at all (three mbarrier waits per stage, usually already complete); the fp64 validation mode keeps
// its predicted state in global memory and synchronises with three cluster barriers per stage instead.
//
// Plain C++ shared between nvcc and g++ like rbc2d_core.h: tests/emu runs the CTAs of a cluster one after
// the other inside every phase, with "remote" pointers resolved inside one host arena.
#pragma once
#include "rbc2d_core.h"

#if defined(__CUDACC__)
#include <cooperative_groups.h>
#endif

namespace rbc2dx {

using rbc2d::BoolTag;
using rbc2d::Consts;
using rbc2d::EnvIO;
using rbc2d::HostConfig;
using rbc2d::HostWrappers;
using rbc2d::RunFlags;
using rbc2d::cx;
using rbc2d::cadd;
using rbc2d::csub;
using rbc2d::cmul;
using rbc2d::muli;
using rbc2d::ldc;
using rbc2d::stc;
using rbc2d::upwind5;
using rbc2d::upwind3;
using rbc2d::upwind1;
using rbc2d::upwind_ord;
using rbc2d::centred4;
using rbc2d::centred_ord;
using rbc2d::untangle_pair;
using rbc2d::tangle_pair;

// ------------------------------------------------------------------------------------------
// compile-time description of one registered grid and its decomposition
// ------------------------------------------------------------------------------------------
template <int NX_, int NZ_, int CL_, int NSTRIP_, int N1_ = 6>
struct Grid {
    static constexpr int NX = NX_, NZ = NZ_, CL = CL_;
    static constexpr int NZL = NZ / CL;                  // rows of the slab one CTA owns
    static constexpr int HALO = 3;
    static constexpr int NSTRIP = NSTRIP_, RS = NZL / NSTRIP;
    static constexpr int NT = NX * NSTRIP;               // threads per CTA
    static constexpr int SX = NX + 2;                    // padded on-chip row (SX mod 32 == 2, see rbc2d_core.h)
    static constexpr int LR = NZL + 2 * HALO;            // on-chip rows per field; local row lr <-> global k0 - 3 + lr
    static constexpr int OFF_B = 0, OFF_U = LR * SX, OFF_W = 2 * LR * SX;
    static constexpr int NS_SM = 3 * LR * SX;            // words per on-chip state buffer
    static constexpr int NCELL = NX * NZ, NWF = NX * (NZ + 1), NSTATE = 2 * NCELL + NWF;
    static constexpr int GOFF_B = 0, GOFF_U = NCELL, GOFF_W = 2 * NCELL;
    static constexpr int NLOC = 3 * NZL * NX;            // tendencies one CTA keeps per stage
    static constexpr int N1 = N1_, NH = NX / 2, N2 = NH / N1;   // NH-point complex FFT = N1 x N2 Cooley-Tukey
    static constexpr int RSTR = SX;
    static constexpr int NR = (NZL + 1) * RSTR;          // Poisson scratch: NZL rows + the pressure row below the slab
    static constexpr int NRED = 11;                      // partial sums per thread in the epilogue
    static constexpr int NFIN = 16;                      // per-CTA totals exchanged through DSMEM
    static constexpr int NEDGE = NX / 32, NE = 3 * NZL * NEDGE;   // left fluxes of the warps' first columns (phase_edge_fluxes)
    static_assert(NZ % CL == 0 && NZL % NSTRIP == 0, "slabs and strips must tile the grid");
    static_assert(RS >= 4, "a strip must be at least 4 rows (full-order fluxes at strip boundaries)");
    static_assert(NZL % 8 == 0, "two-sided Thomas sweeps the block in chunks of 4 rows from both ends");
    static_assert(NX % 32 == 0 && SX % 32 == 2, "rows must map whole warps and pad to 2 banks");
    static_assert(NH == N1 * N2 && (N1 == 4 || N1 == 6) && (N2 == 4 || N2 == 8 || N2 == 16), "unsupported FFT split");
    static_assert(NT <= 1024, "too many threads");
};

// everything one CTA needs while its cluster owns an environment.  Shared-memory regions are byte offsets
// from the CTA's base so that the host emulator can keep one arena per cluster rank.
template <typename Real>
struct CtxX {
    unsigned char* base;        // device: this CTA's dynamic shared memory; host: arena of rank 0
    size_t arena_stride;        // host: bytes between the arenas of consecutive ranks (device: unused)
    unsigned o_s0, o_s1, o_R, o_Tb, o_mid, o_ends, o_twN, o_tw2, o_red, o_fin, o_cfin, o_bars, o_tinv, o_phs, o_edge;   // o_R / o_tinv == ~0u: see project(); o_phs: split mode only
    Real* gm;                   // global: [CL][2][NLOC] stage tendencies of this cluster (ping-pong slabs)
    Real* nxt_g;                // global: [CL][NS_SM] predicted state (fp64 mode) or nullptr
    // global tables, [CL] blocks each (build_tables_host)
    const Real* tinv;           // [CL][NZL][NX] two-sided Thomas pivots of the local block
    const Real* spv;            // [CL][NZL][NX] spike  A_j^-1 e_first
    const Real* spw;            // [CL][NZL][NX] spike  A_j^-1 e_last
    const Real* cxl;            // [CL][2 CL][NX] rows of the inverted reduced system: x_{j-1}[last]
    const Real* cxr;            // [CL][2 CL][NX] ... and x_{j+1}[first]
    const Real* twN;            // [NH][2]  cos/sin(2 pi j / NH)
    const Real* tw2;            // [NH][2]  cos/sin(2 pi m / NX)
    Real thomas_scale;          // dz^2 / NH
};

#if defined(__CUDACC__)
#define RBX_RANK() ((int)cooperative_groups::this_cluster().block_rank())
#define RBX_SYNC_LOCAL __syncthreads()
#define RBX_SYNC_CLUSTER(G) do { if (G::CL > 1) cooperative_groups::this_cluster().sync(); else __syncthreads(); } while (0)
#endif
#if defined(__CUDA_ARCH__)
template <typename T>
__device__ __forceinline__ T* peer_ptr(const unsigned char*, size_t, T* p, int /*my*/, int to)
{
    return cooperative_groups::this_cluster().map_shared_rank(p, to);
}
#define RBX_PHASE_L(G, ...) { const int tid = threadIdx.x; const int rank = my_rank; unsigned char* const smb = X.base; (void)rank; (void)smb; __VA_ARGS__ } RBX_SYNC_LOCAL;
#define RBX_PHASE_C(G, ...) { const int tid = threadIdx.x; const int rank = my_rank; unsigned char* const smb = X.base; (void)rank; (void)smb; __VA_ARGS__ } RBX_SYNC_CLUSTER(G);
// a phase whose cross-CTA stores are tracked by the receivers' mbarriers (ASYNCF) needs only the CTA barrier
#define RBX_PHASE_X(G, ASYNCF, ...) { const int tid = threadIdx.x; const int rank = my_rank; unsigned char* const smb = X.base; (void)rank; (void)smb; __VA_ARGS__ } \
    do { if (ASYNCF) __syncthreads(); else RBX_SYNC_CLUSTER(G); } while (0);
#else
template <typename T>
inline T* peer_ptr(const unsigned char*, size_t stride, T* p, int my, int to)
{
    return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(p) + (ptrdiff_t)(to - my) * (ptrdiff_t)stride);
}
#define RBX_PHASE_L(G, ...) for (int rank = 0; rank < G::CL; ++rank) { unsigned char* const smb = X.base + rank * X.arena_stride; (void)smb; \
        for (int tid = 0; tid < G::NT; ++tid) { __VA_ARGS__ } }
#define RBX_PHASE_C(G, ...) RBX_PHASE_L(G, __VA_ARGS__)
#define RBX_PHASE_X(G, ASYNCF, ...) RBX_PHASE_L(G, __VA_ARGS__)
#endif

// ------------------------------------------------------------------------------------------
// cross-CTA stores.  A PeerBuf names a buffer of another CTA of the cluster:
//   host emulator      : a pointer into that rank's arena;
//   device, barrier mode: a generic pointer (DSMEM through map_shared_rank, or the peer's global scratch);
//   device, async mode  : the shared::cluster address of the buffer and of the receiver's mbarrier; every
//                         store is a st.async that completes sizeof(Real) bytes on that mbarrier.
// Channels (one pair of mbarriers each, alternating with the use count so that consecutive uses never share
// a barrier): W = the w* face from the slab above, E = block-end values of the tridiagonal solve, H = halo rows,
// P = column totals of the hydrostatic pressure integral from the slabs above (split mode only).
// ------------------------------------------------------------------------------------------
//
// Why no cluster barrier is needed (stage s, buffers ping-pong A/B, barriers alternate with the use count):
//   * a CTA starts the tendency of stage s+1 only after wait H(s), i.e. after BOTH neighbours finished their
//     correct(s); it reaches thomas_back(s) only after its own tendency(s); it passes wait E(s) only after ALL
//     CTAs ran thomas_back(s).  Hence, when a CTA pushes b*/w* in tendency(s) into a neighbour's `nxt` halos,
//     that neighbour has passed thomas_back(s-1) (we waited E(s-1)) and no longer reads that buffer (it was its
//     `cur` in tendency(s-1)); when a CTA pushes corrected u,w (and overwrites the shared w* face) in correct(s),
//     the neighbour has passed thomas_back(s) (we waited E(s)) and therefore its pass A of stage s.
//   * ends[parity] slot of rank r in CTA q is rewritten by r in thomas_back(s+2); r got there through wait E(s+1),
//     so q ran thomas_back(s+1), which follows its spike_correct(s) — the last reader of the slot.
//   * barrier (ch, parity) is re-used by pushes of use n+2; a producer two uses ahead would have needed the
//     consumer's own pushes of use n+1, which the consumer issues only after completing its wait of use n.
enum { CH_W = 0, CH_E = 1, CH_H = 2, CH_P = 3, NCHAN = 4 };   // CH_P: column sums of the hydrostatic pressure (split mode)
struct SyncState { unsigned n[NCHAN]; };              // uses of each channel so far (identical on all CTAs)

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned a, unsigned r)
{
    unsigned o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
    return o;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void st_async(unsigned addr, float v, unsigned bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async(unsigned addr, double v, unsigned bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr), "l"(__double_as_longlong(v)), "r"(bar) : "memory");
}
#endif

// Stress build (-DRBX_JITTER, librbc_b200_jitter.so): pseudo-random delays in front of every remote store and every mbarrier
// wait, different per CTA, warp and run (clock-seeded) and biased against one rank of the cluster at a time, so that the
// CTAs of a cluster drift apart as far as the protocol lets them.  The results must stay bitwise identical to the normal
// build (tests/test_gpu_grid192.py): the write-after-read argument above is then tested, not only argued.
#if defined(__CUDA_ARCH__) && defined(RBX_JITTER)
__device__ __forceinline__ void rbx_jitter(unsigned site)
{
    unsigned h = ((unsigned)clock() * 2654435761u) ^ (blockIdx.x * 40503u) ^ (site * 0x9E3779B1u) ^ ((threadIdx.x >> 5) * 7919u);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    const unsigned slow_rank = ((unsigned)(clock64() >> 16)) & 7u;               // one rank is the laggard for ~30 us at a time
    if ((h & 15u) == 0u) __nanosleep((h >> 8) & 0x3FFu);
    if ((blockIdx.x & 7u) == slow_rank && (h & 3u) == 1u) __nanosleep(0x200u + ((h >> 20) & 0x3FFu));
}
#define RBX_JITTER_AT(site) rbx_jitter(site)
#else
#define RBX_JITTER_AT(site) do { } while (0)
#endif

template <typename Real>
struct PeerBuf {
    Real* p;            // nullptr: there is no such peer
    unsigned addr, bar; // async mode only
};
// peer `to`'s copy of the shared-memory region at byte offset data_off; stores complete on its barrier at bar_off
template <bool ASYNC, typename Real>
RBC_HD PeerBuf<Real> make_peer(const CtxX<Real>& X, unsigned char* smb, unsigned data_off, unsigned bar_off, int my, int to, bool exists)
{
    PeerBuf<Real> pb;
    pb.p = nullptr; pb.addr = 0; pb.bar = 0;
    if (!exists) return pb;
    Real* lp = reinterpret_cast<Real*>(smb + data_off);
#if defined(__CUDA_ARCH__)
    if (ASYNC) {
        pb.addr = mapa_u32(smem_u32(lp), (unsigned)to);
        pb.bar = mapa_u32(smem_u32(smb + bar_off), (unsigned)to);
        pb.p = lp;                                           // non-null marker only
    } else {
        pb.p = cooperative_groups::this_cluster().map_shared_rank(lp, to);
    }
    (void)my; (void)X;
#else
    (void)bar_off;
    pb.p = peer_ptr(X.base, X.arena_stride, lp, my, to);
#endif
    return pb;
}
template <bool ASYNC, typename Real>
RBC_HD void peer_store(const PeerBuf<Real>& pb, int idx, Real v)
{
#if defined(__CUDA_ARCH__)
    if (ASYNC) { RBX_JITTER_AT(1u); st_async(pb.addr + (unsigned)idx * (unsigned)sizeof(Real), v, pb.bar); }
    else pb.p[idx] = v;
#else
    pb.p[idx] = v;
#endif
}
// byte offset of the mbarrier serving use number n of channel ch
RBC_HD unsigned bar_off(unsigned o_bars, int ch, unsigned n) { return o_bars + (unsigned)(ch * 2 + (int)(n & 1u)) * 8u; }

#if defined(__CUDA_ARCH__)
#define RBX_ASYNC(G, NXTG) (G::CL > 1 && !(NXTG))
// consumer side: thread 0 arms the barrier with the bytes this CTA is about to receive, everybody waits for them
#define RBX_WAIT(ch, cnt, bytes)                                                        \
    do {                                                                                \
        const unsigned nbytes_ = (unsigned)(bytes);                                     \
        if (nbytes_ > 0) {                                                              \
            const unsigned bar_ = smem_u32(X.base + bar_off(X.o_bars, (ch), (cnt)));    \
            RBX_JITTER_AT(2u + (unsigned)(ch));                                         \
            if (threadIdx.x == 0) mbar_expect_tx(bar_, nbytes_);                        \
            mbar_wait(bar_, ((cnt) >> 1) & 1u);                                         \
        }                                                                               \
    } while (0)
#else
#define RBX_ASYNC(G, NXTG) false
#define RBX_WAIT(ch, cnt, bytes) do { (void)(bytes); } while (0)
#endif

template <typename G> RBC_HD int wrapx(int i) { return i < 0 ? i + G::NX : (i >= G::NX ? i - G::NX : i); }

// wall-order rules (0-based GLOBAL indices), SURVEY 8a
template <typename G> RBC_HD int ord_up_face(int kf) { return (kf >= 3 && kf <= G::NZ - 3) ? 5 : ((kf == 2 || kf == G::NZ - 2) ? 3 : 1); }
template <typename G> RBC_HD int ord_ce_face(int kf) { return (kf >= 2 && kf <= G::NZ - 2) ? 4 : 2; }
template <typename G> RBC_HD int ord_up_cen(int kc) { return (kc >= 2 && kc <= G::NZ - 3) ? 5 : ((kc == 1 || kc == G::NZ - 2) ? 3 : 1); }
template <typename G> RBC_HD int ord_ce_cen(int kc) { return (kc >= 1 && kc <= G::NZ - 2) ? 4 : 2; }

// ------------------------------------------------------------------------------------------
// phase: load the slab (with halos) of one environment from the checkpoint layout in global memory
// ------------------------------------------------------------------------------------------
template <typename G, typename Real, typename Src>
RBC_HD void phase_load_state(int tid, int rank, const Src* RBC_RESTRICT g, Real* RBC_RESTRICT sm)
{
    const int k0 = rank * G::NZL - G::HALO;
    for (int q = tid; q < 3 * G::LR * G::NX; q += G::NT) {
        const int f = q / (G::LR * G::NX), rem = q % (G::LR * G::NX), lr = rem / G::NX, i = rem % G::NX;
        const int k = k0 + lr;
        const int kmax = (f == 2) ? G::NZ : G::NZ - 1;
        const int goff = (f == 0) ? G::GOFF_B : (f == 1 ? G::GOFF_U : G::GOFF_W);
        sm[f * G::LR * G::SX + lr * G::SX + i] = (k >= 0 && k <= kmax) ? (Real)g[goff + k * G::NX + i] : Real(0);
    }
}
template <typename G, typename Real>
RBC_HD void phase_store_state(int tid, int rank, const Real* RBC_RESTRICT sm, Real* RBC_RESTRICT g)
{
    const int k0 = rank * G::NZL;
    for (int q = tid; q < 3 * G::NZL * G::NX; q += G::NT) {
        const int f = q / (G::NZL * G::NX), rem = q % (G::NZL * G::NX), lk = rem / G::NX, i = rem % G::NX;
        const int goff = (f == 0) ? G::GOFF_B : (f == 1 ? G::GOFF_U : G::GOFF_W);
        g[goff + (k0 + lk) * G::NX + i] = sm[f * G::LR * G::SX + (lk + G::HALO) * G::SX + i];
    }
    if (rank == G::CL - 1)                                   // top wall face of w
        for (int i = tid; i < G::NX; i += G::NT) g[G::GOFF_W + G::NZ * G::NX + i] = Real(0);
}

// ------------------------------------------------------------------------------------------
// shared x-fluxes (see rbc2d_core.h left_fluxes / phase_edge_fluxes): every thread evaluates the three LEFT fluxes of
// its cell and takes the right ones from lane+1; lane 31 reads those of the next warp's first column from a table
// filled by this prologue.  Local row lk of the slab, on-chip row lk + HALO; the u window may reach into the halo rows
// (zero beyond the walls, like every window of the march).
// ------------------------------------------------------------------------------------------
template <typename G, typename Real>
RBC_HD void left_fluxes(const Real* RBC_RESTRICT c, int i, int lk, int kg, bool want_bu, bool want_w, Real& fb, Real& fu, Real& fw)
{
    constexpr int SX = G::SX, H = G::HALO;
    const Real* RBC_RESTRICT cb = c + G::OFF_B + (lk + H) * SX;
    const Real* RBC_RESTRICT cu = c + G::OFF_U + (lk + H) * SX;
    const Real* RBC_RESTRICT cw = c + G::OFF_W + (lk + H) * SX;
    int col[6];
    RBC_UNROLL
    for (int j = 0; j < 6; ++j) col[j] = wrapx<G>(i - 3 + j);
    if (want_bu) {
        Real bx[6], ux[6];
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) { bx[j] = cb[col[j]]; ux[j] = cu[col[j]]; }
        fb = upwind5(ux[3], bx);
        fu = upwind5(centred4(ux[1], ux[2], ux[3], ux[4]), ux);
    }
    if (want_w) {
        Real wx[6], uz[4];
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) wx[j] = cw[col[j]];
        RBC_UNROLL
        for (int j = 0; j < 4; ++j) uz[j] = c[G::OFF_U + (lk + H - 2 + j) * SX + i];
        fw = upwind5(centred_ord(uz[0], uz[1], uz[2], uz[3], ord_ce_face<G>(kg)), wx);
    }
}
template <typename G, typename Real>
RBC_HD void phase_edge_fluxes(int tid, int rank, const Real* RBC_RESTRICT c, Real* RBC_RESTRICT E)
{
    constexpr int NZL = G::NZL, NEDGE = G::NEDGE;
    for (int q = tid; q < 2 * NZL * NEDGE; q += G::NT) {
        const int item = q % (NZL * NEDGE), half = q / (NZL * NEDGE);
        const int lk = item % NZL, ci = item / NZL;
        Real fb = Real(0), fu = Real(0), fw = Real(0);
        left_fluxes<G>(c, ci * 32, lk, rank * NZL + lk, half == 0, half == 1, fb, fu, fw);
        if (half == 0) {
            E[(0 * NZL + lk) * NEDGE + ci] = fb;
            E[(1 * NZL + lk) * NEDGE + ci] = fu;
        } else {
            E[(2 * NZL + lk) * NEDGE + ci] = fw;
        }
    }
}

// ------------------------------------------------------------------------------------------
// phase: tendencies + RK3 substep of one strip of the slab (see rbc2d_core.h phase_tendency; same
// arithmetic in the same order).  Halo rows make every window load unconditional; b* rows and the w* face
// that the neighbouring slabs need before the next barrier are pushed into their buffers as they are made.
// ------------------------------------------------------------------------------------------
template <typename G, bool ASYNC, bool SPLIT, typename Real>
RBC_HD void phase_tendency(int tid, int rank, const Consts<Real>& C, const Real* RBC_RESTRICT c, Real* RBC_RESTRICT n,
                           const PeerBuf<Real>& below_h, const PeerBuf<Real>& below_w, const PeerBuf<Real>& above_h,
                           const Real* RBC_RESTRICT phy, const Real* RBC_RESTRICT Tb, const Real* RBC_RESTRICT E, const Real* gm_in, Real* gm_out,
                           Real dt, Real gam, Real zet, bool use_gm)
{
    constexpr int NX = G::NX, NZ = G::NZ, SX = G::SX, RS = G::RS, NT = G::NT, NZL = G::NZL, H = G::HALO, NEDGE = G::NEDGE;
    const int i = tid % NX, s = tid / NX, lk0 = s * RS, kg0 = rank * NZL + lk0;
    const bool last_lane = (i & 31) == 31;
    const int eci = ((i + 1) % NX) / 32;                 // table column of lane 31's right neighbour
    const Real* RBC_RESTRICT cb = c + G::OFF_B;
    const Real* RBC_RESTRICT cu = c + G::OFF_U;
    const Real* RBC_RESTRICT cw = c + G::OFF_W;
    int col[6];
    RBC_UNROLL
    for (int j = 0; j < 6; ++j) col[j] = wrapx<G>(i - 3 + j);

    // sliding windows of the own column; index j <-> local row lk0 - 3 + j (on-chip row lk0 + j)
    Real bz[7], uz[7], wz[7], wxr[6];
    RBC_UNROLL
    for (int j = 0; j < 7; ++j) {
        bz[j] = cb[(lk0 + j) * SX + i];
        uz[j] = cu[(lk0 + j) * SX + i];
        wz[j] = cw[(lk0 + j) * SX + i];
    }
    RBC_UNROLL
    for (int j = 0; j < 6; ++j) wxr[j] = cw[(lk0 + H) * SX + col[j]];            // w(i-3..i+2, face k0)

    Real Fzb_lo = Real(0), Wu_lo = Real(0), Ww_lo = Real(0);
    if (kg0 >= 1) {                                       // strip boundaries are interior faces: full order
        Fzb_lo = upwind5(wz[3], bz);
        Wu_lo = upwind5(centred4(wxr[1], wxr[2], wxr[3], wxr[4]), uz);
        Ww_lo = upwind5(centred4(wz[1], wz[2], wz[3], wz[4]), wz);
    }
    const Real tb = Tb[i];
    const Real kdx = C.kappa * C.idx2, kdz = C.kappa * C.idz2, ndx = C.nu * C.idx2, ndz = C.nu * C.idz2;
    const Real dtg = dt * gam, dtz = dt * zet;

    Real gnb = Real(0), gnu = Real(0), gnw = Real(0);
    if (use_gm) { gnb = gm_in[(0 * RS) * NT + tid]; gnu = gm_in[(1 * RS) * NT + tid]; gnw = gm_in[(2 * RS) * NT + tid]; }

    auto row = [&](auto edge_tag, const int r) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int lk = lk0 + r, k = kg0 + r, lr = lk + H;
        const Real gb0 = gnb, gu0 = gnu, gw0 = gnw;
        if (r + 1 < RS && use_gm) {
            gnb = gm_in[(0 * RS + r + 1) * NT + tid];
            gnu = gm_in[(1 * RS + r + 1) * NT + tid];
            gnw = gm_in[(2 * RS + r + 1) * NT + tid];
        }
        Real bx[6], ux[6], wxn[6];
        RBC_UNROLL
        for (int j = 0; j < 6; ++j) {
            bx[j] = (j == 3) ? bz[3] : cb[lr * SX + col[j]];
            ux[j] = (j == 3) ? uz[3] : cu[lr * SX + col[j]];
            wxn[j] = (j == 3) ? wz[4] : cw[(lr + 1) * SX + col[j]];
        }
        const bool top = EDGE && (k == NZ - 1);
        const bool bot = EDGE && (k == 0);
        const int o_face_hi = EDGE ? ord_up_face<G>(k + 1) : 5;
        const int o_ce_face = EDGE ? ord_ce_face<G>(k) : 4;
        const int o_up_cen = EDGE ? ord_up_cen<G>(k) : 5;
        const int o_ce_cen = EDGE ? ord_ce_cen<G>(k) : 4;

        // left x-fluxes; the right ones are the neighbour's left ones
        const Real Fx0 = upwind5(ux[3], bx);
        const Real F0 = upwind5(centred4(ux[1], ux[2], ux[3], ux[4]), ux);
        const Real Fw0 = upwind5(centred_ord(uz[1], uz[2], uz[3], uz[4], o_ce_face), wxr);
        Real Fx1, F1, Fw1;
#if defined(__CUDA_ARCH__)
        Fx1 = __shfl_down_sync(0xffffffffu, Fx0, 1);
        F1 = __shfl_down_sync(0xffffffffu, F0, 1);
        Fw1 = __shfl_down_sync(0xffffffffu, Fw0, 1);
        if (last_lane) {
            Fx1 = E[(0 * NZL + lk) * NEDGE + eci];
            F1 = E[(1 * NZL + lk) * NEDGE + eci];
            Fw1 = E[(2 * NZL + lk) * NEDGE + eci];
        }
#else
        if (last_lane) {
            Fx1 = E[(0 * NZL + lk) * NEDGE + eci];
            F1 = E[(1 * NZL + lk) * NEDGE + eci];
            Fw1 = E[(2 * NZL + lk) * NEDGE + eci];
        } else {
            left_fluxes<G>(c, i + 1, lk, k, true, true, Fx1, F1, Fw1);
        }
#endif

        // ---- tracer ----
        const Real Fzb_hi = top ? Real(0) : upwind_ord(wz[4], bz + 1, o_face_hi);
        const Real bdn = bot ? (Real(2) * tb - bz[3]) : bz[2];
        const Real bup = top ? (Real(2) * C.b_top - bz[3]) : bz[4];
        const Real Gb = (Fx0 - Fx1) * C.idx + (Fzb_lo - Fzb_hi) * C.idz + (bx[4] - Real(2) * bx[3] + bx[2]) * kdx +
                        (bup - Real(2) * bz[3] + bdn) * kdz;

        // ---- u ----
        const Real Wu_hi = top ? Real(0) : upwind_ord(centred4(wxn[1], wxn[2], wxn[3], wxn[4]), uz + 1, o_face_hi);
        const Real udn = bot ? -uz[3] : uz[2];
        const Real uup = top ? -uz[3] : uz[4];
        Real Gu = (F0 - F1) * C.idx + (Wu_lo - Wu_hi) * C.idz + (ux[4] - Real(2) * ux[3] + ux[2]) * ndx +
                  (uup - Real(2) * uz[3] + udn) * ndz;
        if (SPLIT) Gu -= (phy[lk * G::RSTR + i] - phy[lk * G::RSTR + col[2]]) * C.idx;

        // ---- w (face k; face 0 is the wall) ----
        const Real Ww_hi = upwind_ord(centred_ord(wz[2], wz[3], wz[4], wz[5], o_ce_cen), wz + 1, o_up_cen);
        Real Gw = (Fw0 - Fw1) * C.idx + (Ww_lo - Ww_hi) * C.idz + (wxr[4] - Real(2) * wxr[3] + wxr[2]) * ndx +
                  (wz[4] - Real(2) * wz[3] + wz[2]) * ndz;
        if (!SPLIT) Gw += Real(0.5) * (bz[2] + bz[3]);
        if (bot) Gw = Real(0);

        // ---- RK3 substep ----
        gm_out[(0 * RS + r) * NT + tid] = Gb;
        gm_out[(1 * RS + r) * NT + tid] = Gu;
        gm_out[(2 * RS + r) * NT + tid] = Gw;
        const Real bn = bz[3] + dtg * Gb + dtz * gb0;
        const Real wn = bot ? Real(0) : wz[3] + dtg * Gw + dtz * gw0;
        n[G::OFF_B + lr * SX + i] = bn;
        n[G::OFF_U + lr * SX + i] = uz[3] + dtg * Gu + dtz * gu0;
        n[G::OFF_W + lr * SX + i] = wn;
        // halo pushes (b is final after this phase; the w* face shared with the slab below feeds its divergence)
        if (G::CL > 1) {
            if (lk < H && below_h.p != nullptr) {
                peer_store<ASYNC>(below_h, G::OFF_B + (NZL + H + lk) * SX + i, bn);
                if (lk == 0) peer_store<ASYNC>(below_w, G::OFF_W + (NZL + H) * SX + i, wn);
            }
            if (lk >= NZL - H && above_h.p != nullptr) peer_store<ASYNC>(above_h, G::OFF_B + (lk - NZL + H) * SX + i, bn);
        }

        // ---- slide ----
        Fzb_lo = Fzb_hi; Wu_lo = Wu_hi; Ww_lo = Ww_hi;
        if (r + 1 < RS) {
            RBC_UNROLL
            for (int j = 0; j < 6; ++j) { bz[j] = bz[j + 1]; uz[j] = uz[j + 1]; wz[j] = wz[j + 1]; }
            bz[6] = cb[(lr + 4) * SX + i];
            uz[6] = cu[(lr + 4) * SX + i];
            wz[6] = cw[(lr + 4) * SX + i];
            RBC_UNROLL
            for (int j = 0; j < 6; ++j) wxr[j] = wxn[j];
        }
    };

    RBC_UNROLL
    for (int r = 0; r < RS; ++r) {
        if (r < 2 || r >= RS - 3) {                         // rows that touch a wall in the bottom / top strip of the domain
            const int k = kg0 + r;
            if (k < 2 || k >= NZ - 3) row(BoolTag<true>{}, r);
            else row(BoolTag<false>{}, r);
        } else {
            row(BoolTag<false>{}, r);
        }
    }
}

template <typename Real>
RBC_HD void phase_copy(int tid, int nt, const Real* src, Real* dst, int nvals)
{
    for (int q = tid; q < nvals; q += nt) dst[q] = src[q];
}

// ------------------------------------------------------------------------------------------
// split mode (pressure channels requested): the hydrostatic pressure anomaly pHY' of the slab — the column
// integral of b from the top wall, as in rbc2d_core.h phase_phy.  Every CTA integrates its own rows starting
// from zero and sends its column totals to the slabs below; adding the totals of the slabs above then gives
// the absolute values.  `phy` is the Poisson scratch (free during the tendency), `tot` is [CL][NX].
// ------------------------------------------------------------------------------------------
template <typename G, bool ASYNC, typename Real>
RBC_HD void phase_phy_local(int tid, int rank, const Consts<Real>& C, const Real* RBC_RESTRICT cb, Real* RBC_RESTRICT phy,
                            const PeerBuf<Real>* peers)
{
    constexpr int NX = G::NX, NZL = G::NZL, SX = G::SX, H = G::HALO, RSTR = G::RSTR;
    if (tid >= NX) return;
    const Real dz = Real(1) / C.idz;
    Real above = cb[(NZL - 1 + H) * SX + tid];
    const Real over = (rank == G::CL - 1) ? (Real(2) * C.b_top - above) : cb[(NZL + H) * SX + tid];
    Real acc = -Real(0.5) * (above + over) * dz;
    phy[(NZL - 1) * RSTR + tid] = acc;
    for (int lk = NZL - 2; lk >= 0; --lk) {
        const Real bk = cb[(lk + H) * SX + tid];
        acc = acc - Real(0.5) * (bk + above) * dz;
        phy[lk * RSTR + tid] = acc;
        above = bk;
    }
    for (int j = 0; j < rank; ++j) peer_store<ASYNC>(peers[j], rank * NX + tid, acc);
}
template <typename G, typename Real>
RBC_HD void phase_phy_offset(int tid, int rank, const Real* RBC_RESTRICT tot, Real* RBC_RESTRICT phy)
{
    constexpr int NX = G::NX, RS = G::RS, RSTR = G::RSTR;
    if (rank == G::CL - 1) return;
    const int i = tid % NX, s = tid / NX;
    Real off = Real(0);
    for (int j = G::CL - 1; j > rank; --j) off += tot[j * NX + i];
    for (int r = 0; r < RS; ++r) phy[(s * RS + r) * RSTR + i] += off;
}

// ------------------------------------------------------------------------------------------
// small complex DFTs by size (natural order in and out)
// ------------------------------------------------------------------------------------------
template <int SIGN, typename Real>
RBC_HD void dft16(cx<Real>* x)
{
    cx<Real> e[8], o[8];
    RBC_UNROLL
    for (int j = 0; j < 8; ++j) { e[j] = x[2 * j]; o[j] = x[2 * j + 1]; }
    rbc2d::dft8<SIGN>(e);
    rbc2d::dft8<SIGN>(o);
    const Real c1 = Real(0.92387953251128675613), s1 = Real(0.38268343236508977173), h = Real(0.70710678118654752440);
    // W16^k = cos(pi k/8) + SIGN i sin(pi k/8)
    const cx<Real> w1 = cmul(o[1], c1, Real(SIGN) * s1);
    const cx<Real> w2 = cmul(o[2], h, Real(SIGN) * h);
    const cx<Real> w3 = cmul(o[3], s1, Real(SIGN) * c1);
    const cx<Real> w4 = muli<SIGN>(o[4]);
    const cx<Real> w5 = cmul(o[5], -s1, Real(SIGN) * c1);
    const cx<Real> w6 = cmul(o[6], -h, Real(SIGN) * h);
    const cx<Real> w7 = cmul(o[7], -c1, Real(SIGN) * s1);
    x[0] = cadd(e[0], o[0]); x[8] = csub(e[0], o[0]);
    x[1] = cadd(e[1], w1); x[9] = csub(e[1], w1);
    x[2] = cadd(e[2], w2); x[10] = csub(e[2], w2);
    x[3] = cadd(e[3], w3); x[11] = csub(e[3], w3);
    x[4] = cadd(e[4], w4); x[12] = csub(e[4], w4);
    x[5] = cadd(e[5], w5); x[13] = csub(e[5], w5);
    x[6] = cadd(e[6], w6); x[14] = csub(e[6], w6);
    x[7] = cadd(e[7], w7); x[15] = csub(e[7], w7);
}
template <int SIGN, int N, typename Real>
RBC_HD void dftn(cx<Real>* x)
{
    if (N == 4) rbc2d::dft4<SIGN>(x[0], x[1], x[2], x[3]);
    else if (N == 6) rbc2d::dft6<SIGN>(x);
    else if (N == 8) rbc2d::dft8<SIGN>(x);
    else dft16<SIGN>(x);
}

// ------------------------------------------------------------------------------------------
// The x-transform of the slab rows (see rbc2d_core.h for the 6 x 8 case).  n = n1 + N1 n2, k = N2 k1 + k2.
//   pass A fwd : divergence on the fly -> N2-point DFTs over n2 -> twiddle W_NH^(n1 k2) -> slot n1 + N1 k2
//   pass B fwd : N1-point DFTs over n1 for the column pair (k2, N2-k2) [or (0, N2/2)] -> Z[N2 k1 + k2] in
//                slot k1 + N1 k2, real-FFT split in registers -> 2 real spectral words per slot
//   pass B inv / pass A inv : exact inverses (unnormalised).
// item = group * ROWS + row: lanes walk consecutive rows (conflict-free 64-bit accesses on padded rows).
// ------------------------------------------------------------------------------------------
template <typename G, typename Real>
RBC_HD void fft_passA_fwd_div(int item, const Consts<Real>& C, const Real* RBC_RESTRICT p, Real* RBC_RESTRICT R,
                              const Real* RBC_RESTRICT twN)
{
    constexpr int N1 = G::N1, N2 = G::N2, SX = G::SX, H = G::HALO;
    const int n1 = item / G::NZL, row = item % G::NZL;
    const Real* pu = p + G::OFF_U + (row + H) * SX;
    const Real* pw0 = p + G::OFF_W + (row + H) * SX;
    const Real* pw1 = pw0 + SX;
    cx<Real> a[N2];
    RBC_UNROLL
    for (int n2 = 0; n2 < N2; ++n2) {
        const int x = 2 * (n1 + N1 * n2);
        const cx<Real> u01 = ldc(pu + x), w0 = ldc(pw0 + x), w1 = ldc(pw1 + x);
        const Real u2 = pu[(x + 2 == G::NX) ? 0 : x + 2];
        a[n2].re = (u01.im - u01.re) * C.idx + (w1.re - w0.re) * C.idz;
        a[n2].im = (u2 - u01.im) * C.idx + (w1.im - w0.im) * C.idz;
    }
    dftn<-1, N2>(a);
    Real* z = R + row * G::RSTR;
    RBC_UNROLL
    for (int k2 = 0; k2 < N2; ++k2) {
        const int j = n1 * k2;
        stc(z + 2 * (n1 + N1 * k2), cmul(a[k2], twN[2 * j], -twN[2 * j + 1]));
    }
}
template <typename G, typename Real>
RBC_HD void fft_passA_inv(int n1, Real* z)
{
    constexpr int N1 = G::N1, N2 = G::N2;
    cx<Real> a[N2];
    RBC_UNROLL
    for (int k2 = 0; k2 < N2; ++k2) a[k2] = ldc(z + 2 * (n1 + N1 * k2));
    dftn<+1, N2>(a);
    RBC_UNROLL
    for (int n2 = 0; n2 < N2; ++n2) stc(z + 2 * (n1 + N1 * n2), a[n2]);
}
template <typename G, typename Real>
RBC_HD void fft_passB_fwd_untangle(int g, Real* RBC_RESTRICT z, const Real* RBC_RESTRICT tw2)
{
    constexpr int N1 = G::N1, N2 = G::N2;
    const int ka = g, kb = (g == 0) ? N2 / 2 : N2 - g;
    Real* za = z + 2 * N1 * ka;
    Real* zb = z + 2 * N1 * kb;
    cx<Real> a[N1], b[N1];
    RBC_UNROLL
    for (int n1 = 0; n1 < N1; ++n1) { a[n1] = ldc(za + 2 * n1); b[n1] = ldc(zb + 2 * n1); }
    dftn<-1, N1>(a);                                  // a[k1] = Z[N2 k1 + ka]
    dftn<-1, N1>(b);
    if (g == 0) {
        const Real x0 = a[0].re + a[0].im, xh = a[0].re - a[0].im;
        a[0] = {x0, xh};                              // (X[0], X[NH]) share slot 0
        RBC_UNROLL
        for (int k1 = 1; k1 < N1 / 2; ++k1) untangle_pair(a[k1], a[N1 - k1], tw2[2 * (N2 * k1)], tw2[2 * (N2 * k1) + 1]);
        a[N1 / 2].im = -a[N1 / 2].im;                 // m = NH/2 is its own partner
        RBC_UNROLL
        for (int k1 = 0; k1 < N1 / 2; ++k1) {
            const int m = N2 * k1 + N2 / 2;
            untangle_pair(b[k1], b[N1 - 1 - k1], tw2[2 * m], tw2[2 * m + 1]);
        }
    } else {
        RBC_UNROLL
        for (int k1 = 0; k1 < N1; ++k1) {
            const int m = N2 * k1 + g;
            untangle_pair(a[k1], b[N1 - 1 - k1], tw2[2 * m], tw2[2 * m + 1]);
        }
    }
    RBC_UNROLL
    for (int k1 = 0; k1 < N1; ++k1) { stc(za + 2 * k1, a[k1]); stc(zb + 2 * k1, b[k1]); }
}
template <typename G, typename Real>
RBC_HD void fft_passB_inv_tangle(int g, Real* RBC_RESTRICT z, const Real* RBC_RESTRICT twN, const Real* RBC_RESTRICT tw2)
{
    constexpr int N1 = G::N1, N2 = G::N2;
    const int ka = g, kb = (g == 0) ? N2 / 2 : N2 - g;
    Real* za = z + 2 * N1 * ka;
    Real* zb = z + 2 * N1 * kb;
    cx<Real> a[N1], b[N1];
    RBC_UNROLL
    for (int k1 = 0; k1 < N1; ++k1) { a[k1] = ldc(za + 2 * k1); b[k1] = ldc(zb + 2 * k1); }
    if (g == 0) {
        const Real x0 = a[0].re, xh = a[0].im;
        a[0] = {Real(0.5) * (x0 + xh), Real(0.5) * (x0 - xh)};
        RBC_UNROLL
        for (int k1 = 1; k1 < N1 / 2; ++k1) tangle_pair(a[k1], a[N1 - k1], tw2[2 * (N2 * k1)], tw2[2 * (N2 * k1) + 1]);
        a[N1 / 2].im = -a[N1 / 2].im;
        RBC_UNROLL
        for (int k1 = 0; k1 < N1 / 2; ++k1) {
            const int m = N2 * k1 + N2 / 2;
            tangle_pair(b[k1], b[N1 - 1 - k1], tw2[2 * m], tw2[2 * m + 1]);
        }
    } else {
        RBC_UNROLL
        for (int k1 = 0; k1 < N1; ++k1) {
            const int m = N2 * k1 + g;
            tangle_pair(a[k1], b[N1 - 1 - k1], tw2[2 * m], tw2[2 * m + 1]);
        }
    }
    dftn<+1, N1>(a);
    dftn<+1, N1>(b);
    RBC_UNROLL
    for (int n1 = 0; n1 < N1; ++n1) {
        const int ja = n1 * ka, jb = n1 * kb;
        stc(za + 2 * n1, cmul(a[n1], twN[2 * ja], twN[2 * ja + 1]));
        stc(zb + 2 * n1, cmul(b[n1], twN[2 * jb], twN[2 * jb + 1]));
    }
}

// ------------------------------------------------------------------------------------------
// z-tridiagonal solve, SPIKE-partitioned over the cluster.  Local block of this CTA (rows 0..M-1, M = NZL):
//   x(k-1) + D_k x(k) + x(k+1) = r(k),  with x(-1) = xl (last value of the block below), x(M) = xr
// (1) phase_thomas_sweep / phase_thomas_back: two-sided Thomas on the block with xl = xr = 0 -> y
//     (pivot tables as in rbc2d_core.h, per cluster rank), block-end values y(0), y(M-1) -> `ends`;
// (2) after a cluster barrier, phase_spike_correct: xl, xr from the pre-inverted reduced system applied to
//     the 2 CL end values of all blocks, then x = y - v xl - w xr;  xl itself is the (spectral) pressure row
//     below the slab and is stored as row M of the scratch.
// ------------------------------------------------------------------------------------------
template <typename G, typename Real>
RBC_HD void phase_thomas_sweep(int tid, Real* RBC_RESTRICT R, const Real* RBC_RESTRICT tinv, Real* RBC_RESTRICT mid, Real scale)
{
    constexpr int NX = G::NX, M = G::NZL, MH = M / 2, RSTR = G::RSTR;
    if (tid >= 2 * NX) return;
    const int t = tid % NX;
    const bool hi = tid >= NX;
    constexpr int BK = (MH % 8 == 0) ? 8 : 4;
    const int sgn = hi ? -1 : 1;
    Real* RBC_RESTRICT pr = R + (hi ? (M - 1) * RSTR : 0) + t;
    const Real* RBC_RESTRICT pt = tinv + (hi ? (M - 1) * NX : 0) + t;
    const int sr = sgn * RSTR, st = sgn * NX;
    Real d = Real(0);
    for (int kb = 0; kb < MH; kb += BK) {
        Real iv[BK], rs[BK];
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) { iv[j] = pt[(kb + j) * st]; rs[j] = pr[(kb + j) * sr]; }
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) rs[j] = rs[j] * scale * iv[j];
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) { d = rs[j] - d * iv[j]; pr[(kb + j) * sr] = d; }
    }
    mid[tid] = d;
}
template <typename G, bool ASYNC, typename Real>
RBC_HD void phase_thomas_back(int tid, int rank, Real* RBC_RESTRICT R, const Real* RBC_RESTRICT tinv, const Real* RBC_RESTRICT mid,
                              Real* RBC_RESTRICT ends, const PeerBuf<Real>* peers)
{
    constexpr int NX = G::NX, M = G::NZL, MH = M / 2, RSTR = G::RSTR;
    if (tid >= 2 * NX) return;
    const int t = tid % NX;
    const bool hi = tid >= NX;
    constexpr int BK = (MH % 8 == 0) ? 8 : 4;
    const Real dlo = mid[t], ehi = mid[NX + t], ivm = tinv[(MH - 1) * NX + t], jvm = tinv[MH * NX + t];
    const Real pl = (dlo - ivm * ehi) / (Real(1) - ivm * jvm);
    const Real ph = ehi - jvm * pl;
    Real pv = hi ? ph : pl;
    const int sgn = hi ? 1 : -1;
    Real* RBC_RESTRICT pr = R + (hi ? MH : MH - 1) * RSTR + t;
    const Real* RBC_RESTRICT pt = tinv + (hi ? MH : MH - 1) * NX + t;
    const int sr = sgn * RSTR, st = sgn * NX;
    pr[0] = pv;
    for (int kb = 1; kb < MH; kb += BK) {
        Real iv[BK], dd[BK];
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) {
            const bool ok = kb + j < MH;
            iv[j] = ok ? pt[(kb + j) * st] : Real(0);
            dd[j] = ok ? pr[(kb + j) * sr] : Real(0);
        }
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) {
            if (kb + j < MH) {
                pv = dd[j] - iv[j] * pv;
                pr[(kb + j) * sr] = pv;
            }
        }
    }
    // block-end values y(0) (slot t) and y(M-1) (slot NX + t) of this rank, to every CTA of the cluster
    ends[rank * 2 * NX + tid] = pv;
    if (G::CL > 1) {
        RBC_UNROLL
        for (int j = 0; j < G::CL; ++j)
            if (j != rank) peer_store<ASYNC>(peers[j], rank * 2 * NX + tid, pv);
    }
}
template <typename G, typename Real>
RBC_HD void phase_spike_correct(int tid, int rank, const CtxX<Real>& X, const Real* RBC_RESTRICT ends, Real* RBC_RESTRICT R)
{
    constexpr int NX = G::NX, M = G::NZL, MH = M / 2, RSTR = G::RSTR, CL = G::CL;
    if (tid >= 2 * NX) return;
    const int t = tid % NX;
    const bool hi = tid >= NX;
    const Real* cl = X.cxl + (size_t)rank * 2 * CL * NX + t;
    const Real* cr = X.cxr + (size_t)rank * 2 * CL * NX + t;
    Real xl = Real(0), xr = Real(0);
    RBC_UNROLL
    for (int j = 0; j < CL; ++j) {
        const Real e0 = ends[j * 2 * NX + t], e1 = ends[j * 2 * NX + NX + t];
        xl += cl[(2 * j) * NX] * e0 + cl[(2 * j + 1) * NX] * e1;
        xr += cr[(2 * j) * NX] * e0 + cr[(2 * j + 1) * NX] * e1;
    }
    const int kbeg = hi ? MH : 0;
    const Real* RBC_RESTRICT pv = X.spv + ((size_t)rank * M + kbeg) * NX + t;
    const Real* RBC_RESTRICT pw = X.spw + ((size_t)rank * M + kbeg) * NX + t;
    Real* RBC_RESTRICT pr = R + kbeg * RSTR + t;
    constexpr int BK = (MH % 8 == 0) ? 8 : 4;
    for (int kb = 0; kb < MH; kb += BK) {
        Real v[BK], w[BK], y[BK];
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) { v[j] = pv[(kb + j) * NX]; w[j] = pw[(kb + j) * NX]; y[j] = pr[(kb + j) * RSTR]; }
        RBC_UNROLL
        for (int j = 0; j < BK; ++j) pr[(kb + j) * RSTR] = y[j] - v[j] * xl - w[j] * xr;
    }
    if (!hi) R[M * RSTR + t] = xl;                    // spectral pressure row just below the slab
}

// ------------------------------------------------------------------------------------------
// phase: pressure correction of the slab + push of the corrected halo rows to the neighbouring slabs
// ------------------------------------------------------------------------------------------
template <typename G, bool ASYNC, typename Real>
RBC_HD void phase_correct(int tid, int rank, const Consts<Real>& C, Real* RBC_RESTRICT p, const PeerBuf<Real>& below, const PeerBuf<Real>& above,
                          const Real* RBC_RESTRICT R)
{
    constexpr int NX = G::NX, SX = G::SX, RS = G::RS, RSTR = G::RSTR, NZL = G::NZL, H = G::HALO;
    const int i = tid % NX, s = tid / NX, im = wrapx<G>(i - 1), lk0 = s * RS, kg0 = rank * NZL + lk0;
    Real* RBC_RESTRICT pu = p + G::OFF_U;
    Real* RBC_RESTRICT pw = p + G::OFF_W;
    Real ph[RS + 1], pm[RS], uu[RS], ww[RS];
    ph[0] = (lk0 >= 1) ? R[(lk0 - 1) * RSTR + i] : R[NZL * RSTR + i];      // row below the slab: from the reduced solve
    RBC_UNROLL
    for (int r = 0; r < RS; ++r) {
        const int lk = lk0 + r;
        ph[r + 1] = R[lk * RSTR + i];
        pm[r] = R[lk * RSTR + im];
        uu[r] = pu[(lk + H) * SX + i];
        ww[r] = pw[(lk + H) * SX + i];
    }
    RBC_UNROLL
    for (int r = 0; r < RS; ++r) {
        const int lk = lk0 + r, k = kg0 + r;
        const Real un = uu[r] - (ph[r + 1] - pm[r]) * C.idx;
        const Real wn = (k >= 1) ? ww[r] - (ph[r + 1] - ph[r]) * C.idz : ww[r];
        pu[(lk + H) * SX + i] = un;
        if (k >= 1) pw[(lk + H) * SX + i] = wn;
        if (G::CL > 1) {
            if (lk < H && below.p != nullptr) {
                peer_store<ASYNC>(below, G::OFF_U + (NZL + H + lk) * SX + i, un);
                peer_store<ASYNC>(below, G::OFF_W + (NZL + H + lk) * SX + i, wn);
            }
            if (lk >= NZL - H && above.p != nullptr) {
                peer_store<ASYNC>(above, G::OFF_U + (lk - NZL + H) * SX + i, un);
                peer_store<ASYNC>(above, G::OFF_W + (lk - NZL + H) * SX + i, wn);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// RBCRewardShaping.compute_cell_distances on one row of w (see rbc2d_core.h cell_distance), any NX
// ------------------------------------------------------------------------------------------
template <int NX, typename Real>
RBC_HD double cell_distance(const Real* uy)
{
    const double PI = 3.14159265358979323846;
    int peaks[NX / 2 + 1];
    int np = 0;
    int i = 1;
    const int imax = NX - 1;
    while (i < imax) {
        if (uy[i - 1] < uy[i]) {
            int ahead = i + 1;
            while (ahead < imax && uy[ahead] == uy[i]) ++ahead;
            if (uy[ahead] < uy[i]) {
                const int mid = (i + ahead - 1) / 2;
                if ((double)uy[mid] >= 0.001) peaks[np++] = mid;
                i = ahead;
            }
        }
        ++i;
    }
    if (np <= 1) return 0.0;
    double best = 0.0;
    for (int a = 0; a < np; ++a)
        for (int b = a + 1; b < np; ++b) {
            const double xa = peaks[a] * (2 * PI / NX), xb = peaks[b] * (2 * PI / NX);
            const double d1 = fabs(xb - xa), d2 = 2 * PI - d1;
            double d = d1 < d2 ? d1 : d2;
            bool pos = true;
            if (d1 < d2) {
                for (int q = peaks[a]; q < peaks[b]; ++q) pos = pos && (uy[q] > Real(0));
            } else {
                for (int q = peaks[b]; q < NX; ++q) pos = pos && (uy[q] > Real(0));
                for (int q = 0; q < peaks[a]; ++q) pos = pos && (uy[q] > Real(0));
            }
            if (pos) d = 0.0;
            if (d > best) best = d;
        }
    return best;
}

// ------------------------------------------------------------------------------------------
// the Poisson projection of the slab state at byte offset `o_p` (FFT-x, SPIKE tridiagonal-z)
// ------------------------------------------------------------------------------------------
#define RBX_PTR(off) reinterpret_cast<Real*>(smb + (off))

// Thomas pivots of this rank's block: the on-chip copy (fp32 device path) or the global table
template <typename G, typename Real, bool TINV_GLOBAL>
RBC_HD const Real* local_tinv(const CtxX<Real>& X, unsigned char* smb, int rank)
{
#if defined(__CUDA_ARCH__)
    if (!TINV_GLOBAL) return reinterpret_cast<const Real*>(smb + X.o_tinv);
#endif
    (void)smb;
    return X.tinv + (size_t)rank * G::NZL * G::NX;
}

template <typename G, typename Real, bool NXT_GLOBAL, bool SPLIT>
RBC_HD void project(const Consts<Real>& C, const CtxX<Real>& X, unsigned o_p, unsigned o_dead, int my_rank, SyncState& S, bool after_tendency)
{
    (void)my_rank;
    constexpr bool OWN_R = NXT_GLOBAL || SPLIT;            // split mode keeps the scratch (it outlives the projection: pHY', pNHS)
    // Poisson scratch: with two on-chip state buffers the one that is dead during the projection (the old state)
    // provides it — its u rows, which no neighbour writes before this CTA has finished correct() (b*/w* pushes of the
    // next tendency only touch b halo rows and one w row).  The shared memory saved holds this rank's Thomas pivots.
    const unsigned o_R = OWN_R ? X.o_R : o_dead + (unsigned)(G::OFF_U * sizeof(Real));

    constexpr int NZL = G::NZL, NT = G::NT, NX = G::NX, N1 = G::N1, N2 = G::N2, CL = G::CL;
    constexpr bool ASYNC = RBX_ASYNC(G, NXT_GLOBAL);
    constexpr unsigned ROWB = NX * sizeof(Real);           // bytes of one pushed row
    if (ASYNC && after_tendency) {                         // the w* face on top of the slab, pushed by the slab above
        RBX_WAIT(CH_W, S.n[CH_W], (my_rank < CL - 1) ? ROWB : 0u);
    }
    if (after_tendency) S.n[CH_W] += 1;
    RBX_PHASE_L(G,
        for (int item = tid; item < N1 * NZL; item += NT)
            fft_passA_fwd_div<G>(item, C, RBX_PTR(o_p), RBX_PTR(o_R), RBX_PTR(X.o_twN));
    )
    RBX_PHASE_L(G,
        for (int item = tid; item < (N2 / 2) * NZL; item += NT)
            fft_passB_fwd_untangle<G>(item / NZL, RBX_PTR(o_R) + (item % NZL) * G::RSTR, RBX_PTR(X.o_tw2));
    )
    RBX_PHASE_L(G, phase_thomas_sweep<G>(tid, RBX_PTR(o_R), local_tinv<G, Real, OWN_R>(X, smb, rank), RBX_PTR(X.o_mid), X.thomas_scale);)
    const unsigned o_ends = X.o_ends + (S.n[CH_E] & 1u) * (unsigned)(CL * 2 * NX * sizeof(Real));     // double-buffered by use parity
    RBX_PHASE_X(G, ASYNC,
        PeerBuf<Real> peers[CL];
        RBC_UNROLL
        for (int j = 0; j < CL; ++j)
            peers[j] = make_peer<ASYNC, Real>(X, smb, o_ends, bar_off(X.o_bars, CH_E, S.n[CH_E]), rank, j, j != rank);
        phase_thomas_back<G, ASYNC>(tid, rank, RBX_PTR(o_R), local_tinv<G, Real, OWN_R>(X, smb, rank), RBX_PTR(X.o_mid), RBX_PTR(o_ends), peers);
    )
    if (CL > 1) {
        if (ASYNC) { RBX_WAIT(CH_E, S.n[CH_E], (unsigned)(CL - 1) * 2u * ROWB); }
        RBX_PHASE_L(G, phase_spike_correct<G>(tid, rank, X, RBX_PTR(o_ends), RBX_PTR(o_R));)
    }
    S.n[CH_E] += 1;
    // inverse transforms; the extra row NZL (pressure row below the slab) rides on otherwise idle lanes
    RBX_PHASE_L(G,
        constexpr int NB = (N2 / 2) * NZL;
        constexpr int NBX = (CL > 1) ? NB + N2 / 2 : NB;
        for (int item = tid; item < NBX; item += NT) {
            if (item < NB) fft_passB_inv_tangle<G>(item / NZL, RBX_PTR(o_R) + (item % NZL) * G::RSTR, RBX_PTR(X.o_twN), RBX_PTR(X.o_tw2));
            else fft_passB_inv_tangle<G>(item - NB, RBX_PTR(o_R) + NZL * G::RSTR, RBX_PTR(X.o_twN), RBX_PTR(X.o_tw2));
        }
    )
    RBX_PHASE_L(G,
        constexpr int NA = N1 * NZL;
        constexpr int NAX = (CL > 1) ? NA + N1 : NA;
        for (int item = tid; item < NAX; item += NT) {
            if (item < NA) fft_passA_inv<G, Real>(item / NZL, RBX_PTR(o_R) + (item % NZL) * G::RSTR);
            else fft_passA_inv<G, Real>(item - NA, RBX_PTR(o_R) + NZL * G::RSTR);
        }
    )
    RBX_PHASE_X(G, ASYNC,
        const unsigned hb = bar_off(X.o_bars, CH_H, S.n[CH_H]);
        const PeerBuf<Real> below = make_peer<ASYNC, Real>(X, smb, o_p, hb, rank, rank - 1, CL > 1 && rank > 0);
        const PeerBuf<Real> above = make_peer<ASYNC, Real>(X, smb, o_p, hb, rank, rank + 1, CL > 1 && rank < CL - 1);
        phase_correct<G, ASYNC>(tid, rank, C, RBX_PTR(o_p), below, above, RBX_PTR(o_R));
    )
    if (ASYNC) {
        // halo rows for the next tendency: 3 rows each of u and w (this phase) and, after a tendency, of b*, per neighbour
        const unsigned nbrs = (my_rank > 0 ? 1u : 0u) + (my_rank < CL - 1 ? 1u : 0u);
        RBX_WAIT(CH_H, S.n[CH_H], nbrs * (after_tendency ? 9u : 6u) * ROWB);
    }
    S.n[CH_H] += 1;
}

// ------------------------------------------------------------------------------------------
// one action step of one environment, executed by all CTAs of the cluster
// ------------------------------------------------------------------------------------------
// split mode: pHY' of the slab state at byte offset o_state into the Poisson scratch (two phases and one exchange)
template <typename G, typename Real, bool NXT_GLOBAL>
RBC_HD void hydrostatic(const Consts<Real>& C, const CtxX<Real>& X, unsigned o_state, int my_rank, SyncState& S)
{
    (void)my_rank;
    constexpr bool ASYNC = RBX_ASYNC(G, NXT_GLOBAL);
    constexpr int CL = G::CL;
    RBX_PHASE_X(G, ASYNC,
        PeerBuf<Real> peers[CL];
        RBC_UNROLL
        for (int j = 0; j < CL; ++j)
            peers[j] = make_peer<ASYNC, Real>(X, smb, X.o_phs, bar_off(X.o_bars, CH_P, S.n[CH_P]), rank, j, j < rank);
        phase_phy_local<G, ASYNC>(tid, rank, C, RBX_PTR(o_state) + G::OFF_B, RBX_PTR(X.o_R), peers);
    )
    if (CL > 1) {
        if (ASYNC) { RBX_WAIT(CH_P, S.n[CH_P], (unsigned)(CL - 1 - my_rank) * (unsigned)(G::NX * sizeof(Real))); }
        RBX_PHASE_L(G, phase_phy_offset<G>(tid, rank, RBX_PTR(X.o_phs), RBX_PTR(X.o_R));)
    }
    S.n[CH_P] += 1;
}

// ------------------------------------------------------------------------------------------
// CFL guard (fp32 throughput mode).  Above CFL ~ 1.5 the RK3 / 5th-order-upwind scheme is linearly unstable; the first plume
// burst of a high-Ra flow started from noise gets there, fp64 rides it out on its 1e-9 smaller round-off
// seed, fp32 leaves ~1 environment in 1000 with NaNs.  At the top of every fourth RK3 step the cluster measures
// max(|w| dt/dz, |u| dt/dx) of the environment (own rows, warp maxima, one DSMEM exchange behind a cluster barrier — 3 barriers
// per measurement against ~25 k cycles per stage) and, above the limit, takes the step as ceil(CFL) equal parts.  Every CTA computes
// the same number, so the mbarrier channel counts stay in step.  Environments below the limit — all developed flows — are
// untouched bit for bit.  The measurement is repeated every CFL_EVERY RK3 steps (always with the full dt_solver) and the split
// kept in between (`covers` = RK3 steps this measurement stands for, for the event counter).  Returns the number of parts (>= 1).
// ------------------------------------------------------------------------------------------
constexpr int CFL_EVERY = 4;
template <typename G, typename Real>
RBC_HD int cfl_parts(const Consts<Real>& C, const EnvIO<Real>& io, const CtxX<Real>& X, int env, unsigned o_cur, int my_rank, Real dt, int sub, int covers)
{
    (void)my_rank;
    constexpr int NX = G::NX, NT = G::NT, SX = G::SX, H = G::HALO;
    const int slot = 13 + (sub & 1);                       // two cfin slots alternate: a slow peer may still be reading the other one
    RBX_PHASE_L(G, if (tid == 0) reinterpret_cast<double*>(smb + X.o_cfin)[slot] = 0.0;)
    RBX_PHASE_C(G,
        const Real* p = RBX_PTR(o_cur);
        const int i = tid % NX; const int s = tid / NX;
        Real m = Real(0);
        for (int r = 0; r < G::RS; ++r) {
            const int lr = s * G::RS + r + H;
            const Real aw = fabs(p[G::OFF_W + lr * SX + i]) * C.idz; const Real au = fabs(p[G::OFF_U + lr * SX + i]) * C.idx;
            m = aw > m ? aw : m; m = au > m ? au : m;      // a NaN never wins: the environment is lost anyway and reported as such
        }
        double* cell = reinterpret_cast<double*>(smb + X.o_cfin) + slot;
#if defined(__CUDA_ARCH__)
        // non-negative floats order like their bit patterns
        unsigned bits = __float_as_uint((float)m);
        bits = __reduce_max_sync(0xffffffffu, bits);
        if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned*>(cell), bits);
#else
        if ((double)m > *cell) *cell = (double)m;
#endif
    )
    double mx = 0.0;
#if defined(__CUDA_ARCH__)
    {
        double* mine = reinterpret_cast<double*>(X.base + X.o_cfin) + slot;
        for (int j = 0; j < G::CL; ++j) {
            const unsigned b = *reinterpret_cast<const unsigned*>((j == my_rank) ? mine : peer_ptr(X.base, X.arena_stride, mine, my_rank, j));
            const double v = (double)__uint_as_float(b);
            mx = v > mx ? v : mx;
        }
    }
#else
    for (int j = 0; j < G::CL; ++j) {
        const double v = reinterpret_cast<const double*>(X.base + j * X.arena_stride + X.o_cfin)[slot];
        mx = v > mx ? v : mx;
    }
    (void)NT;
#endif
    const double cfl = mx * (double)dt;
    int parts = 1;
    if (cfl > (double)C.cfl_limit) { parts = (int)ceil(cfl); parts = parts < 2 ? 2 : (parts > 8 ? 8 : parts); }
    if (parts > 1 && io.cfl_events != nullptr) {
        RBX_PHASE_L(G, if (rank == 0 && tid == 0) io.cfl_events[env] += (parts - 1) * covers;)
    }
    return parts;
}

// sum over the CTAs of the cluster of entry `idx` of their cfin blocks (valid after the cluster barrier that follows the
// phase writing them); every thread of every CTA gets the same value
template <typename G, typename Real>
RBC_HD double cluster_cfin_sum(const CtxX<Real>& X, int my_rank, int idx)
{
    double tot = 0;
#if defined(__CUDA_ARCH__)
    double* mine = reinterpret_cast<double*>(X.base + X.o_cfin);
    for (int j = 0; j < G::CL; ++j) tot += (j == my_rank) ? mine[idx] : peer_ptr(X.base, X.arena_stride, mine, my_rank, j)[idx];
#else
    (void)my_rank;
    for (int j = 0; j < G::CL; ++j) tot += reinterpret_cast<const double*>(X.base + j * X.arena_stride + X.o_cfin)[idx];
#endif
    return tot;
}

// epilogue of an action step (see rbc2d::env_epilogue): runs once per environment visit, and again (second_pass) after an
// in-launch re-initialisation from the checkpoint bank.  Returns whether such a reset is due.
template <typename G, typename Real, bool NXT_GLOBAL, bool SPLIT>
RBC_HD bool env_epilogue(const Consts<Real>& C, const EnvIO<Real>& io, const CtxX<Real>& X, int env, const RunFlags& F, int my_rank,
                   SyncState& S, unsigned o_cur, unsigned o_nxt, Real last_dtau, bool state_changed, bool second_pass)
{
    (void)my_rank;
    constexpr int NX = G::NX, NZ = G::NZ, NZL = G::NZL, NT = G::NT, CL = G::CL, SX = G::SX, H = G::HALO, NRED = G::NRED, NFIN = G::NFIN;
    const rbc2d::VecIO& V = io.vec;
    Real* st = io.state + (size_t)env * G::NSTATE;
    // re-read here rather than carried in registers across the march (see rbc2d::env_epilogue)
    const int pend = (!second_pass && V.mode == 1 && V.bank != nullptr && F.advance_clock) ? V.pending[env] : 0;
    const double t_old = io.t[env];
    state_changed = state_changed || pend;
    const int oz = NZ / C.obs_nz, ox = NX / C.obs_nx, nobs = C.obs_nz * C.obs_nx;
    const unsigned o_red = (X.o_red != ~0u) ? X.o_red : o_nxt;   // fp32: the dead state buffer when it is large enough
    if (SPLIT && io.pressure != nullptr && state_changed) {
        // pressure fields of get_state (rbc_sim2D_api.jl:114-115), kept per environment as in rbc2d_core.h:
        // pNHS = phi / dtau of the last stage in the zero-mean gauge (the mean needs the whole cluster), then
        // pHY' recomputed from the final b
        Real* pr = io.pressure + (size_t)env * 2 * G::NCELL;
        RBX_PHASE_L(G,
            double* red = reinterpret_cast<double*>(smb + o_red);
            const Real* R = RBX_PTR(X.o_R);
            double acc = 0;
            for (int q = tid; q < NZL * NX; q += NT) acc += (double)R[(q / NX) * G::RSTR + (q % NX)];
            red[tid] = acc;
        )
        RBX_PHASE_C(G,
            if (tid == 0) {
                const double* red = reinterpret_cast<const double*>(smb + o_red);
                double acc = 0;
                for (int e = 0; e < NT; ++e) acc += red[e];
                reinterpret_cast<double*>(smb + X.o_cfin)[NRED + 1] = acc;
            }
        )
        RBX_PHASE_L(G,
            double* mine = reinterpret_cast<double*>(smb + X.o_cfin);
            double tot = 0;
            for (int j = 0; j < CL; ++j) tot += (j == rank) ? mine[NRED + 1] : peer_ptr(X.base, X.arena_stride, mine, rank, j)[NRED + 1];
            const double mean = tot / (double)G::NCELL;
            const Real* R = RBX_PTR(X.o_R);
            for (int q = tid; q < NZL * NX; q += NT)
                pr[G::NCELL + rank * NZL * NX + q] = (Real)(((double)R[(q / NX) * G::RSTR + (q % NX)] - mean) / (double)last_dtau);
        )
        hydrostatic<G, Real, NXT_GLOBAL>(C, X, o_cur, my_rank, S);
        RBX_PHASE_L(G,
            const Real* R = RBX_PTR(X.o_R);
            for (int q = tid; q < NZL * NX; q += NT) pr[rank * NZL * NX + q] = R[(q / NX) * G::RSTR + (q % NX)];
        )
    }
    RBX_PHASE_L(G,
        // per-thread partial sums over the thread's strip (get_nusselt, rbc_sim2D_api.jl:142-163):
        //  0: sum b w   1: same on the sensor grid   2..5: sum_x b on rows 0, 1, NZ-2, NZ-1
        //  6: NaN count   7..10: sum over sensor columns of b on sensor rows 0, 1, nzo-2, nzo-1
        const Real* p = RBX_PTR(o_cur);
        double* red = reinterpret_cast<double*>(smb + o_red);
        const int i = tid % NX; const int s = tid / NX;
        const bool xs = (i % ox) == 0;
        double v[NRED];
        for (int q = 0; q < NRED; ++q) v[q] = 0;
        for (int r = 0; r < G::RS; ++r) {
            const int lk = s * G::RS + r; const int k = rank * NZL + lk; const int lr = lk + H;
            const double b = (double)p[G::OFF_B + lr * SX + i]; const double u = (double)p[G::OFF_U + lr * SX + i];
            const double w = (double)p[G::OFF_W + lr * SX + i];
            if (b != b || u != u || w != w) v[6] += 1;
            v[0] += b * w;
            const bool zs = (k % oz) == 0;
            if (zs && xs) v[1] += b * w;
            if (k == 0) v[2] += b;
            if (k == 1) v[3] += b;
            if (k == NZ - 2) v[4] += b;
            if (k == NZ - 1) v[5] += b;
            if (xs) {
                if (k == 0) v[7] += b;
                if (k == oz) v[8] += b;
                if (k == (C.obs_nz - 2) * oz) v[9] += b;
                if (k == (C.obs_nz - 1) * oz) v[10] += b;
            }
        }
        for (int q = 0; q < NRED; ++q) red[q * NT + tid] = v[q];
    )
    RBX_PHASE_L(G,
        // two-level serial sums: NRED x 16 threads over NT/16 values each, then NRED threads over 16
        double* red = reinterpret_cast<double*>(smb + o_red);
        double* fin = reinterpret_cast<double*>(smb + X.o_fin);
        if (tid < NRED * 16) {
            const int q = tid / 16; const int j = tid % 16;
            double acc = 0;
            for (int e = j * (NT / 16); e < (j + 1) * (NT / 16); ++e) acc += red[q * NT + e];
            fin[tid] = acc;
        }
    )
    RBX_PHASE_C(G,
        double* fin = reinterpret_cast<double*>(smb + X.o_fin);
        double* cfin = reinterpret_cast<double*>(smb + X.o_cfin);
        if (tid < NRED) {
            double acc = 0;
            for (int j = 0; j < 16; ++j) acc += fin[tid * 16 + j];
            cfin[tid] = acc;
        }
        if (tid == NRED) {                                   // Benard-cell distance on the mid-height row of w
            double cd = 0.0;
            constexpr int kmid = NZ / 2 - 1;
            if ((C.wrap_shaping || io.cell_dist != nullptr) && kmid / NZL == rank)
                cd = cell_distance<NX>(RBX_PTR(o_cur) + G::OFF_W + (kmid % NZL + H) * SX);
            cfin[NRED] = cd;
        }
    )
    // decisions every thread of every CTA takes alike (cfin blocks are complete: cluster barrier above)
    const bool vec = V.mode >= 0 && F.advance_clock;
    const bool stepping = F.advance_clock && !pend && !second_pass;
    const bool trunc_now = stepping && (t_old + C.dt_action >= C.episode_length);
    const bool bad_reset = vec && stepping && V.nan_reset && cluster_cfin_sum<G>(X, my_rank, 6) > 0;
    const bool do_reset = vec && stepping && V.bank != nullptr && ((V.mode == 2 && trunc_now) || bad_reset);
    float* const ob_base = ((do_reset && V.final_obs != nullptr) ? V.final_obs : io.obs) + (size_t)env * C.channels * nobs;
    RBX_PHASE_C(G,
        const Real* cur = RBX_PTR(o_cur);
        if (state_changed && !do_reset) phase_store_state<G>(tid, rank, cur, st);
        float* ob = ob_base;
        for (int q = tid; q < 3 * nobs; q += NT) {
            const int ch = q / nobs; const int zo = (q % nobs) / C.obs_nx; const int xo = q % C.obs_nx;
            const int k = zo * oz;
            if (k / NZL != rank) continue;
            const int off = (ch == 0 ? G::OFF_B : (ch == 1 ? G::OFF_U : G::OFF_W));
            float v = (float)cur[off + (k % NZL + H) * SX + xo * ox];
            if (C.wrap_obs) {
                v = C.obs_maxval * (2.0f * (v - C.obs_lo[ch]) / (C.obs_hi[ch] - C.obs_lo[ch]) - 1.0f);
                if (C.obs_clip) v = fminf(fmaxf(v, -C.obs_maxval), C.obs_maxval);
            }
            ob[q] = v;
        }
        if (SPLIT && io.pressure != nullptr && C.channels == 5) {     // also on observe-only launches, from the stored fields
            const Real* pr = io.pressure + (size_t)env * 2 * G::NCELL;
            for (int q = tid; q < 2 * nobs; q += NT) {
                const int ch = q / nobs; const int zo = (q % nobs) / C.obs_nx; const int xo = q % C.obs_nx;
                const int k = zo * oz;
                if (k / NZL != rank) continue;
                ob[3 * nobs + q] = (float)pr[(size_t)ch * G::NCELL + k * NX + xo * ox];
            }
        }
        if (rank == 0 && tid == 0) {
            double* my = reinterpret_cast<double*>(smb + X.o_cfin);
            double tot[NFIN];
            for (int q = 0; q <= NRED; ++q) tot[q] = 0;
            for (int j = 0; j < CL; ++j) {
                const double* pf = (j == 0) ? my : peer_ptr(X.base, X.arena_stride, my, 0, j);
                for (int q = 0; q <= NRED; ++q) tot[q] += pf[q];
            }
            const double kap = C.kappa_d; const double dbH = kap * 1.0 / 2.0;
            const double q1 = tot[0] / (double)G::NCELL; const double q1o = tot[1] / (double)nobs;
            const double T0 = tot[2] / NX; const double T1 = tot[3] / NX; const double Tm2 = tot[4] / NX; const double Tm1 = tot[5] / NX;
            const double g = (1.5 * Tm1 - 0.5 * Tm2 + 0.5 * T1 - 1.5 * T0) / NZ;
            const double nu_s = (q1 - kap * g) / dbH;
            const double o0 = tot[7] / C.obs_nx; const double o1 = tot[8] / C.obs_nx; const double om2 = tot[9] / C.obs_nx; const double om1 = tot[10] / C.obs_nx;
            const double go = (1.5 * om1 - 0.5 * om2 + 0.5 * o1 - 1.5 * o0) / C.obs_nz;
            const double nu_o = (q1o - kap * go) / dbH;
            double rew = -nu_o;
            if (C.wrap_reward) rew = (rew + C.reward_scale) / (C.reward_scale - 1.0);
            if (C.wrap_shaping || io.cell_dist != nullptr) {
                const double PI = 3.14159265358979323846;
                const double cd = tot[NRED];
                if (io.cell_dist != nullptr && !second_pass) io.cell_dist[env] = cd;
                if (C.wrap_shaping) rew = (1.0 - C.shaping_weight) * rew + C.shaping_weight * ((PI - cd) / PI);
            }
            const bool bad = tot[6] > 0;
            if (second_pass) {
                io.nu_state[env] = nu_s;
                io.nu_obs[env] = nu_o;
                io.t[env] = 0.0;
                io.step_count[env] = 1;
                V.episode[env] += 1;
                V.ep_return[env] = 0.0;
                if (V.pending != nullptr) V.pending[env] = 0;
            } else if (pend) {
                io.nan_flag[env] = bad ? 1 : 0;
                io.nu_state[env] = nu_s;
                io.nu_obs[env] = nu_o;
                io.reward[env] = 0.0f;
                io.truncated[env] = 0;
                io.t[env] = 0.0;
                io.step_count[env] = 1;
                V.episode[env] += 1;
                V.ep_return[env] = 0.0;
                V.pending[env] = 0;
            } else {
                io.nan_flag[env] = bad ? 1 : 0;
                if (bad_reset) rew = 0.0;
                if (bad && vec && V.nan_count != nullptr) RBC_COUNT_ONE(V.nan_count);
                io.reward[env] = (float)rew;
                const double ret = vec ? V.ep_return[env] + (double)(float)rew : 0.0;
                if (do_reset) {
                    if (V.final_nu_a != nullptr) V.final_nu_a[env] = nu_s;
                    if (V.final_nu_b != nullptr) V.final_nu_b[env] = nu_o;
                    if (V.final_return != nullptr) V.final_return[env] = ret;
                    io.truncated[env] = 1;
                } else {
                    io.nu_state[env] = nu_s;
                    io.nu_obs[env] = nu_o;
                    if (F.advance_clock) {
                        const double tn = t_old + C.dt_action;
                        io.t[env] = tn;
                        io.step_count[env] += 1;
                        io.truncated[env] = (trunc_now || bad_reset) ? 1 : 0;
                    }
                    if (vec) {
                        V.ep_return[env] = ret;
                        if (V.pending != nullptr) V.pending[env] = (V.mode == 1 && trunc_now && V.bank != nullptr) ? 1 : 0;
                    }
                }
            }
        }
    )
    return do_reset;
}

template <typename G, typename Real, bool NXT_GLOBAL, bool SPLIT = false>
RBC_HD void env_action_step(const Consts<Real>& C, const EnvIO<Real>& io, const CtxX<Real>& X, int env, const RunFlags& F, int my_rank,
                            SyncState& S)
{
    (void)my_rank;
    constexpr bool ASYNC = RBX_ASYNC(G, NXT_GLOBAL);
    constexpr int NX = G::NX, NZ = G::NZ, NZL = G::NZL, NT = G::NT, CL = G::CL, SX = G::SX, H = G::HALO, NRED = G::NRED, NFIN = G::NFIN;
    const Real gam[3] = {Real(8.0 / 15.0), Real(5.0 / 12.0), Real(3.0 / 4.0)};
    const Real zet[3] = {Real(0), Real(-17.0 / 60.0), Real(-5.0 / 12.0)};
    Real* st = io.state + (size_t)env * G::NSTATE;
    const rbc2d::VecIO& V = io.vec;
    // fused vector-env semantics (rbc2d_core.h): a pending environment is only re-initialised from the checkpoint bank
    const int pend = (V.mode == 1 && V.bank != nullptr && F.advance_clock) ? V.pending[env] : 0;
    const int nsub = pend ? 0 : F.nsub;
    const bool project_first = pend ? SPLIT : (F.project_first != 0);

    if (pend) {
        const double* src = V.bank + (size_t)rbc2d::checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * G::NSTATE;
        RBX_PHASE_L(G, phase_load_state<G>(tid, rank, src, RBX_PTR(X.o_s0));)
    } else {
        RBX_PHASE_L(G,
            phase_load_state<G>(tid, rank, st, RBX_PTR(X.o_s0));
            if (tid < NX) RBX_PTR(X.o_Tb)[tid] = (Real)rbc2d::heater_T(C, io.actions + (size_t)env * C.heaters, (tid + 0.5) * C.dx);
        )
    }
    unsigned o_cur = X.o_s0, o_nxt = X.o_s1;               // fp64 mode: o_s1 is unused, the predicted state is global
    Real last_dtau = Real(1);                              // set! projects with dtau = 1
    if (project_first) project<G, Real, NXT_GLOBAL, SPLIT>(C, X, o_cur, o_nxt, my_rank, S, false);
    int parts = 1;
    for (int sub = 0; sub < nsub; ++sub) {
        const Real dt_sub = (sub == nsub - 1) ? C.dt_last : C.dt_full;
        // the flow changes little over CFL_EVERY RK3 steps (0.06 time units at dt_solver 0.015): measure once, keep the split
        if (C.cfl_limit > Real(0) && sub % CFL_EVERY == 0) parts = cfl_parts<G>(C, io, X, env, o_cur, my_rank, C.dt_full, sub / CFL_EVERY, nsub - sub < CFL_EVERY ? nsub - sub : CFL_EVERY);
        const Real dt = parts > 1 ? dt_sub / Real(parts) : dt_sub;
        RBC_NOUNROLL
        for (int part = 0; part < parts; ++part)
        for (int stage = 0; stage < 3; ++stage) {
            const int in_slab = (stage & 1) ? 0 : 1, out_slab = 1 - in_slab;
            if (SPLIT) hydrostatic<G, Real, NXT_GLOBAL>(C, X, o_cur, my_rank, S);
            RBX_PHASE_L(G, phase_edge_fluxes<G>(tid, rank, RBX_PTR(o_cur), RBX_PTR(X.o_edge));)
            RBX_PHASE_X(G, ASYNC,
                Real* cur = RBX_PTR(o_cur);
                Real* nxt;
                PeerBuf<Real> below_h; PeerBuf<Real> below_w; PeerBuf<Real> above_h;
                if (NXT_GLOBAL) {
                    nxt = X.nxt_g + (size_t)rank * G::NS_SM;
                    below_h.p = rank > 0 ? nxt - G::NS_SM : nullptr; below_h.addr = 0; below_h.bar = 0;
                    above_h.p = rank < CL - 1 ? nxt + G::NS_SM : nullptr; above_h.addr = 0; above_h.bar = 0;
                    below_w = below_h;
                } else {
                    nxt = RBX_PTR(o_nxt);
                    const unsigned hb = bar_off(X.o_bars, CH_H, S.n[CH_H]);
                    const unsigned wb = bar_off(X.o_bars, CH_W, S.n[CH_W]);
                    below_h = make_peer<ASYNC, Real>(X, smb, o_nxt, hb, rank, rank - 1, CL > 1 && rank > 0);
                    below_w = make_peer<ASYNC, Real>(X, smb, o_nxt, wb, rank, rank - 1, CL > 1 && rank > 0);
                    above_h = make_peer<ASYNC, Real>(X, smb, o_nxt, hb, rank, rank + 1, CL > 1 && rank < CL - 1);
                }
                Real* gmr = X.gm + (size_t)rank * 2 * G::NLOC;
                phase_tendency<G, ASYNC, SPLIT>(tid, rank, C, cur, nxt, below_h, below_w, above_h, SPLIT ? RBX_PTR(X.o_R) : nullptr, RBX_PTR(X.o_Tb), RBX_PTR(X.o_edge), gmr + in_slab * G::NLOC,
                                         gmr + out_slab * G::NLOC, dt, gam[stage], zet[stage], stage > 0);
            )
            unsigned o_p;
            if (NXT_GLOBAL) {
                RBX_PHASE_L(G, phase_copy(tid, NT, X.nxt_g + (size_t)rank * G::NS_SM, RBX_PTR(o_cur), G::NS_SM);)
                o_p = o_cur;
            } else {
                o_p = o_nxt; o_nxt = o_cur; o_cur = o_p;
            }
            project<G, Real, NXT_GLOBAL, SPLIT>(C, X, o_p, NXT_GLOBAL ? o_p : o_nxt, my_rank, S, true);
            last_dtau = (gam[stage] + zet[stage]) * dt;
        }
    }

    // one inlined copy of the epilogue, run a second time when the environment is re-initialised inside this launch
    bool changed = nsub > 0 || project_first;
    RBC_NOUNROLL
    for (int pass = 0; pass < 2; ++pass) {
        if (!env_epilogue<G, Real, NXT_GLOBAL, SPLIT>(C, io, X, env, F, my_rank, S, o_cur, o_nxt, last_dtau, changed, pass == 1)) break;
        const double* src = V.bank + (size_t)rbc2d::checkpoint_draw(V.seed, V.id_offset + (unsigned long long)env, (unsigned long long)V.episode[env], V.n_ep) * G::NSTATE;
        RBX_PHASE_L(G, phase_load_state<G>(tid, rank, src, RBX_PTR(o_cur));)
        if (SPLIT) project<G, Real, NXT_GLOBAL, SPLIT>(C, X, o_cur, o_nxt, my_rank, S, false);
        changed = true; last_dtau = Real(1);
    }
}

// ------------------------------------------------------------------------------------------
// host-side construction of constants and tables (fp64), shared by the library and the emulator
// ------------------------------------------------------------------------------------------
template <typename G, typename Real>
inline Consts<Real> make_consts(const HostConfig& h, const HostWrappers& w = HostWrappers())
{
    Consts<Real> C;
    const double dx = h.lx / G::NX, dz = h.lz / G::NZ;
    const double nu = sqrt(h.pr / h.ra), kappa = 1.0 / sqrt(h.pr * h.ra);
    C.idx = (Real)(1.0 / dx); C.idz = (Real)(1.0 / dz);
    C.idx2 = (Real)(1.0 / (dx * dx)); C.idz2 = (Real)(1.0 / (dz * dz));
    C.nu = (Real)nu; C.kappa = (Real)kappa; C.b_top = (Real)h.b_top;
    double last;
    C.nsub = rbc2d::substep_schedule(h.dt_action, h.dt_solver, &last);
    C.dt_full = (Real)h.dt_solver; C.dt_last = (Real)last;
    C.heaters = h.heaters; C.heater_limit = h.heater_limit; C.lx = h.lx; C.dx = dx;
    C.kappa_d = kappa; C.dt_action = h.dt_action; C.episode_length = h.episode_length;
    C.obs_nz = h.obs_nz; C.obs_nx = h.obs_nx; C.channels = h.channels;
    C.wrap_obs = w.normalize_obs; C.obs_clip = w.obs_clip; C.obs_maxval = w.obs_maxval;
    for (int c = 0; c < 4; ++c) { C.obs_lo[c] = w.obs_lo[c]; C.obs_hi[c] = w.obs_hi[c]; }
    C.wrap_reward = w.normalize_reward; C.reward_scale = w.reward_scale;
    C.wrap_shaping = w.shaping; C.shaping_weight = w.shaping_weight;
    C.cfl_limit = (Real)h.cfl_limit;
    return C;
}

// mode index of spectral word t of a row after fft_passB_fwd_untangle
template <typename G>
inline int word_mode(int t)
{
    const int p = t / 2;
    if (p == 0) return (t == 0) ? 0 : G::NH;
    return G::N2 * (p % G::N1) + p / G::N1;
}

struct HostTables {
    // all fp64; sizes in the comments of CtxX
    double *tinv, *spv, *spw, *cxl, *cxr, *twN, *tw2;
};
template <typename G>
inline size_t table_doubles() { return (size_t)3 * G::NZ * G::NX + (size_t)2 * G::CL * 2 * G::CL * G::NX + 4 * (size_t)G::NH; }

// dense solve of a small system (Gauss-Jordan with partial pivoting) -> inverse, n <= 16
inline void invert_small(int n, const double* A, double* inv)
{
    double a[16][32];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { a[i][j] = A[i * n + j]; a[i][n + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 2 * n; ++j) { const double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double d = 1.0 / a[c][c];
        for (int j = 0; j < 2 * n; ++j) a[c][j] *= d;
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
            if (f != 0.0) for (int j = 0; j < 2 * n; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) inv[i * n + j] = a[i][n + j];
}

template <typename G>
inline void build_tables_host(double lx, double lz, const HostTables& T)
{
    constexpr int NX = G::NX, NZ = G::NZ, CL = G::CL, M = G::NZL, MH = M / 2, NH = G::NH;
    static_assert(2 * CL <= 16, "reduced system too large");
    const double PI = 3.14159265358979323846;
    const double dx = lx / NX, dz = lz / NZ;
    for (int t = 0; t < NX; ++t) {
        const int m = word_mode<G>(t);
        const double sx = 2.0 * sin(PI * m / NX) / dx, lam = sx * sx * dz * dz;
        auto diag = [&](int k) {                           // global row k
            double dg = -(2.0 + lam);
            if (k == 0 || k == NZ - 1) dg += 1.0;
            if (m == 0 && k == 0) dg -= 1.0;               // pin the null space of the mean mode
            return dg;
        };
        double vfirst[CL], vlast[CL], wfirst[CL], wlast[CL];
        for (int j = 0; j < CL; ++j) {
            const int k0 = j * M;
            // two-sided pivots of the local block
            double prev = 0.0;
            for (int k = 0; k < MH; ++k) { const double iv = 1.0 / (diag(k0 + k) - prev); T.tinv[(size_t)(k0 + k) * NX + t] = iv; prev = iv; }
            prev = 0.0;
            for (int k = M - 1; k >= MH; --k) { const double iv = 1.0 / (diag(k0 + k) - prev); T.tinv[(size_t)(k0 + k) * NX + t] = iv; prev = iv; }
            // spikes v = A_j^-1 e_0, w = A_j^-1 e_{M-1}: one-sided Thomas in fp64
            double cp[M], v[M], w[M];
            for (int which = 0; which < 2; ++which) {
                double* x = which ? w : v;
                double dprev = 0.0, cprev = 0.0;
                for (int k = 0; k < M; ++k) {
                    const double rhs = (which == 0) ? (k == 0 ? 1.0 : 0.0) : (k == M - 1 ? 1.0 : 0.0);
                    const double den = diag(k0 + k) - cprev;
                    cp[k] = 1.0 / den;
                    x[k] = (rhs - dprev) / den;
                    cprev = cp[k]; dprev = x[k];
                }
                for (int k = M - 2; k >= 0; --k) x[k] -= cp[k] * x[k + 1];
            }
            for (int k = 0; k < M; ++k) { T.spv[(size_t)(k0 + k) * NX + t] = v[k]; T.spw[(size_t)(k0 + k) * NX + t] = w[k]; }
            vfirst[j] = v[0]; vlast[j] = v[M - 1]; wfirst[j] = w[0]; wlast[j] = w[M - 1];
        }
        // reduced system on z = (x_0[0], x_0[M-1], x_1[0], ...)
        constexpr int NRS = 2 * CL;
        double S[NRS * NRS], Si[NRS * NRS];
        for (int q = 0; q < NRS * NRS; ++q) S[q] = 0.0;
        for (int j = 0; j < CL; ++j) {
            S[(2 * j) * NRS + 2 * j] = 1.0;
            S[(2 * j + 1) * NRS + 2 * j + 1] = 1.0;
            if (j > 0) { S[(2 * j) * NRS + 2 * j - 1] = vfirst[j]; S[(2 * j + 1) * NRS + 2 * j - 1] = vlast[j]; }
            if (j < CL - 1) { S[(2 * j) * NRS + 2 * j + 2] = wfirst[j]; S[(2 * j + 1) * NRS + 2 * j + 2] = wlast[j]; }
        }
        invert_small(NRS, S, Si);
        for (int j = 0; j < CL; ++j)
            for (int q = 0; q < NRS; ++q) {
                T.cxl[((size_t)j * NRS + q) * NX + t] = (j > 0) ? Si[(2 * j - 1) * NRS + q] : 0.0;
                T.cxr[((size_t)j * NRS + q) * NX + t] = (j < CL - 1) ? Si[(2 * j + 2) * NRS + q] : 0.0;
            }
    }
    for (int j = 0; j < NH; ++j) { T.twN[2 * j] = cos(2 * PI * j / NH); T.twN[2 * j + 1] = sin(2 * PI * j / NH); }
    for (int m = 0; m < NH; ++m) { T.tw2[2 * m] = cos(2 * PI * m / NX); T.tw2[2 * m + 1] = sin(2 * PI * m / NX); }
}

// shared-memory layout of one CTA (bytes); identical on the device and in the host emulator
template <typename G, typename Real, bool NXT_GLOBAL, bool SPLIT = false>
struct SmemLayoutX {
    static constexpr size_t al(size_t x) { return (x + 15) & ~(size_t)15; }
    static constexpr bool OWN_R = NXT_GLOBAL || SPLIT;
    static constexpr size_t s0 = 0;
    static constexpr size_t s1 = al(s0 + sizeof(Real) * G::NS_SM);
    // fp64 (one on-chip state buffer) and split mode: separate Poisson scratch; otherwise the scratch aliases the dead
    // buffer and the space holds this rank's Thomas pivots instead
    static constexpr size_t R = NXT_GLOBAL ? s1 : al(s1 + sizeof(Real) * G::NS_SM);
    static constexpr size_t tinv = R;
    static constexpr size_t red = al(R + (OWN_R ? sizeof(Real) * G::NR : sizeof(Real) * G::NZL * G::NX));
    static_assert(OWN_R || G::LR >= G::NZL + 1, "the scratch must fit the u rows of a state buffer");
    static constexpr bool kRedSeparate = NXT_GLOBAL || sizeof(double) * G::NRED * G::NT > sizeof(Real) * G::NS_SM;
    static constexpr size_t fin = al(red + (kRedSeparate ? sizeof(double) * G::NRED * G::NT : 0));
    static constexpr size_t cfin = al(fin + sizeof(double) * G::NRED * 16);
    static constexpr size_t Tb = al(cfin + sizeof(double) * G::NFIN);
    static constexpr size_t mid = al(Tb + sizeof(Real) * G::NX);
    static constexpr size_t ends = al(mid + sizeof(Real) * 2 * G::NX);
    static constexpr size_t bars = al(ends + sizeof(Real) * 2 * G::CL * 2 * G::NX);     // ends: [2 parities][CL ranks][2 NX]
    // split mode: [CL ranks][NX] column totals of the hydrostatic integral.  With a separate reduction scratch they live
    // inside it, past the NT doubles the pNHS mean uses: the scratch is otherwise idle until the final reductions, which
    // start after the last hydrostatic phase
    static constexpr bool kPhsInRed = SPLIT && kRedSeparate;
    static constexpr size_t phs = kPhsInRed ? red + 4096 : al(bars + 8 * 2 * NCHAN);
    static_assert(!kPhsInRed || (sizeof(double) * G::NT <= 4096 && 4096 + sizeof(Real) * G::CL * G::NX <= sizeof(double) * G::NRED * G::NT),
                  "column totals do not fit the reduction scratch");
    static constexpr size_t twN = al(bars + 8 * 2 * NCHAN + ((SPLIT && !kPhsInRed) ? sizeof(Real) * G::CL * G::NX : 0));
    static constexpr size_t tw2 = al(twN + sizeof(Real) * 2 * G::NH);
    static constexpr size_t edge = al(tw2 + sizeof(Real) * 2 * G::NH);
    static constexpr size_t total = al(edge + sizeof(Real) * G::NE);
    template <typename Ctx>
    static void fill(Ctx& X)
    {
        X.o_s0 = (unsigned)s0; X.o_s1 = (unsigned)s1; X.o_R = OWN_R ? (unsigned)R : ~0u; X.o_tinv = OWN_R ? ~0u : (unsigned)tinv; X.o_red = kRedSeparate ? (unsigned)red : ~0u; X.o_fin = (unsigned)fin;
        X.o_cfin = (unsigned)cfin; X.o_Tb = (unsigned)Tb; X.o_mid = (unsigned)mid; X.o_ends = (unsigned)ends; X.o_bars = (unsigned)bars;
        X.o_phs = SPLIT ? (unsigned)phs : ~0u; X.o_twN = (unsigned)twN; X.o_tw2 = (unsigned)tw2; X.o_edge = (unsigned)edge;
    }
};

}  // namespace rbc2dx
