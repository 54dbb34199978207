/*
 * rbc2d_oracle.c — CPU fp64 restatement of the RBC-Gym 2D simulation step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (rbc_gym_b200/) may link, import
 * or call this file; it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * What it restates (reference = /root/reference, MichielStraat/RBC-Gym):
 *   - step_simulation / run!            src/rbc_gym/sim/rbc_sim2D_api.jl:75-97
 *   - model configuration               src/rbc_gym/sim/rbc_sim2D.jl:149-160
 *       NonhydrostaticModel, UpwindBiasedFifthOrder, :RungeKutta3, BuoyancyTracer,
 *       ScalarDiffusivity(nu,kappa), grid (Periodic, Flat, Bounded) rbc_sim2D.jl:75-84
 *   - boundary conditions / heater      rbc_sim2D.jl:87-146
 *   - get_state / get_observation / get_nusselt / array_gradient
 *                                       rbc_sim2D_api.jl:102-163, rbc_sim2D.jl:206-220
 *   - NaN check                         rbc_sim2D.jl:223-228
 *
 * The PDE arithmetic itself lives in the third-party dependency Oceananigans.jl, pinned at
 * 0.92.0 in src/rbc_gym/juliapkg.json:7-10 and NOT present under /root/reference.  Its
 * published algorithm (staggered C-grid finite volume, 5th-order upwind advection with a
 * 4th-order centred advecting velocity and 5->3->1 / 4->2 order reduction next to walls,
 * explicit diffusion, hydrostatic-pressure split, low-storage RK3 with an FFT/DCT pressure
 * projection per stage) is restated here following SURVEY.md section 8a.
 *
 * PARITY PIN: the reference has no tests and cannot be executed here (no Julia).  The oracle
 * is pinned by the reference's shipped checkpoint files (produced by the real Julia sim):
 * discrete divergence-freeness under this staggering, the Ra=1e4 fixed-point band, the
 * near-wall tracer residual (tests/test_oracle_fixtures.py), the Ra=1e5 limit cycle — all 34
 * two-roll-pair states of the train/val/test files within 0.0075 sigma of this oracle's own
 * trajectory in (Nu_state, KE, Nu_obs) space, the 6 one-roll-pair states inside their branch's
 * band (tests/test_oracle_limit_cycle.py) — plus the survey's cross-check table.  Step-by-step
 * against Julia output itself parity is UNPINNED (no Julia step output exists to compare with).
 *
 * Index conventions (0-based, C order, x fastest):
 *   b[k*nx+i]  cell centre (i,k), k = 0 bottom .. nz-1 top
 *   u[k*nx+i]  x-face on the LEFT of cell i, z-centre k
 *   w[k*nx+i]  z-face BELOW cell k, k = 0 .. nz   (w[0]=w[nz]=0 walls)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    int nx, nz;
    double lx, lz;          /* domain [0,lx] x [0,lz]; reference: 2*pi, 2  (rbc_sim2D_api.jl:28) */
    double nu, kappa;       /* sqrt(Pr/Ra), 1/sqrt(Pr*Ra)            (rbc_sim2D_api.jl:40-41)  */
    double b_top;           /* min_b = 1                              (rbc_sim2D_api.jl:36)     */
    int heaters;            /* 12                                                               */
    double heater_limit;    /* 0.75                                                             */
    int split_phy;          /* 1: buoyancy via hydrostatic pressure anomaly (reference), 0: +b in Gw */
    int poisson_mode;       /* 0: FFT-x + tridiagonal-z, 1: dense DFT-x / DCT-z eigenvalue division   */
} rbc2d_params;

/* ------------------------------------------------------------------------------------- */
/* heater profile: collate_actions_colin + bottom_T, rbc_sim2D.jl:87-138                    */
/* ------------------------------------------------------------------------------------- */
static double heater_T(const rbc2d_params *P, const double *action, double x)
{
    const int na = P->heaters;
    const double ampl = P->heater_limit, dx = 0.03;
    double mean = 0.0;
    for (int s = 0; s < na; ++s) mean += ampl * action[s];
    mean /= na;
    double dev = 0.0;
    for (int s = 0; s < na; ++s) {
        double d = fabs(ampl * action[s] - mean);
        if (d > dev) dev = d;
    }
    double K2 = dev / ampl;
    if (!(K2 > 1.0)) K2 = 1.0;            /* maximum([1.0, ...]) */
    const double seg = P->lx / na;
    int s = (int)floor(x / seg);          /* 0-based segment */
    if (s >= na) s = na - 1;
    const int sm = (s == 0) ? na - 1 : s - 1, sp = (s == na - 1) ? 0 : s + 1;
    const double T0 = 2 + (ampl * action[sm] - mean) / K2;
    const double T1 = 2 + (ampl * action[s] - mean) / K2;
    const double T2 = 2 + (ampl * action[sp] - mean) / K2;
    const double xp = x - s * seg;
    if (xp < dx)
        return T0 + ((T0 - T1) / (4 * dx * dx * dx)) * (xp - 2 * dx) * (xp + dx) * (xp + dx);
    if (xp >= seg - dx)
        return T1 + ((T1 - T2) / (4 * dx * dx * dx)) * (xp - seg - 2 * dx) * (xp - seg + dx) * (xp - seg + dx);
    return T1;
}

void rbc2d_oracle_heater_profile(const rbc2d_params *P, const double *action, double *Tb)
{
    const double dx = P->lx / P->nx;
    for (int i = 0; i < P->nx; ++i) Tb[i] = heater_T(P, action, (i + 0.5) * dx);
}

#include "oracle_stencils.h"

/* ------------------------------------------------------------------------------------- */
/* workspace: x-padded copies (halo 3, periodic) so stencils read straight memory          */
/* ------------------------------------------------------------------------------------- */
#define HX 3
typedef struct {
    int nx, nz, px;           /* px = nx + 2*HX */
    double *bp, *up, *wp;     /* padded fields: b,u: nz rows; w: nz+1 rows */
    double *phy;              /* hydrostatic pressure anomaly, padded like b */
    double *Gb, *Gu, *Gw;     /* tendencies (unpadded)                          */
    double *Gb0, *Gu0, *Gw0;  /* previous-stage tendencies                      */
    double *rhs, *p;          /* nz*nx */
    double *Tb;               /* nx wall temperatures */
    double *fre, *fim;        /* spectral scratch nz*nx each */
    double *cosx, *sinx;      /* twiddles */
} work_t;

static work_t *work_new(int nx, int nz)
{
    work_t *W = (work_t *)calloc(1, sizeof(work_t));
    W->nx = nx; W->nz = nz; W->px = nx + 2 * HX;
    size_t n = (size_t)nx * nz, np = (size_t)W->px * nz, nw = (size_t)nx * (nz + 1), npw = (size_t)W->px * (nz + 1);
    W->bp = calloc(np, 8); W->up = calloc(np, 8); W->wp = calloc(npw, 8); W->phy = calloc(np, 8);
    W->Gb = calloc(n, 8); W->Gu = calloc(n, 8); W->Gw = calloc(nw, 8);
    W->Gb0 = calloc(n, 8); W->Gu0 = calloc(n, 8); W->Gw0 = calloc(nw, 8);
    W->rhs = calloc(n, 8); W->p = calloc(n, 8); W->Tb = calloc(nx, 8);
    W->fre = calloc(n, 8); W->fim = calloc(n, 8);
    W->cosx = calloc(nx, 8); W->sinx = calloc(nx, 8);
    for (int j = 0; j < nx; ++j) { W->cosx[j] = cos(2 * M_PI * j / nx); W->sinx[j] = sin(2 * M_PI * j / nx); }
    return W;
}
static void work_free(work_t *W)
{
    free(W->bp); free(W->up); free(W->wp); free(W->phy); free(W->Gb); free(W->Gu); free(W->Gw);
    free(W->Gb0); free(W->Gu0); free(W->Gw0); free(W->rhs); free(W->p); free(W->Tb);
    free(W->fre); free(W->fim); free(W->cosx); free(W->sinx); free(W);
}

static void pad_rows(double *dst, const double *src, int nx, int rows)
{
    const int px = nx + 2 * HX;
    for (int k = 0; k < rows; ++k) {
        double *d = dst + (size_t)k * px + HX;
        const double *s = src + (size_t)k * nx;
        memcpy(d, s, (size_t)nx * 8);
        for (int h = 1; h <= HX; ++h) { d[-h] = s[nx - h]; d[nx - 1 + h] = s[h - 1]; }
    }
}

/* ------------------------------------------------------------------------------------- */
/* tendencies G(U) — SURVEY 8a bullets Advection / Diffusion / Buoyancy                     */
/* ------------------------------------------------------------------------------------- */
static void tendencies(const rbc2d_params *P, work_t *W, const double *b, const double *u, const double *w)
{
    const int nx = P->nx, nz = P->nz, px = W->px;
    const double dx = P->lx / nx, dz = P->lz / nz;
    const double idx = 1.0 / dx, idz = 1.0 / dz;
    pad_rows(W->bp, b, nx, nz);
    pad_rows(W->up, u, nx, nz);
    pad_rows(W->wp, w, nx, nz + 1);
#define B(i, k) W->bp[(size_t)(k) * px + HX + (i)]
#define U(i, k) W->up[(size_t)(k) * px + HX + (i)]
#define Wv(i, k) W->wp[(size_t)(k) * px + HX + (i)]
#define PH(i, k) W->phy[(size_t)(k) * px + HX + (i)]

    if (P->split_phy) {
        /* hydrostatic pressure anomaly, integrated downward from the top (Oceananigans
         * update_hydrostatic_pressure!): pHY'(Nz) = -0.5(b(Nz)+b(Nz+1)) dz with the top ghost
         * b(Nz+1) = 2 b_top - b(Nz);  pHY'(k) = pHY'(k+1) - 0.5(b(k)+b(k+1)) dz              */
        for (int i = -1; i < nx; ++i) {
            double gh = 2 * P->b_top - B(i, nz - 1);
            PH(i, nz - 1) = -0.5 * (B(i, nz - 1) + gh) * dz;
            for (int k = nz - 2; k >= 0; --k) PH(i, k) = PH(i, k + 1) - 0.5 * (B(i, k) + B(i, k + 1)) * dz;
        }
    }

    for (int k = 0; k < nz; ++k) {
        for (int i = 0; i < nx; ++i) {
            /* ---------------- tracer b at centre (i,k) ---------------- */
            {
                /* x-fluxes at faces i and i+1 */
                double Fx0 = upwind(U(i, k), up5L(&B(i, k), 1), up5R(&B(i, k), 1));
                double Fx1 = upwind(U(i + 1, k), up5L(&B(i + 1, k), 1), up5R(&B(i + 1, k), 1));
                /* z-fluxes at faces k and k+1 (0 at the walls) */
                double Fz0 = 0.0, Fz1 = 0.0;
                if (k >= 1) { int o = ord_up_face(k, nz); Fz0 = upwind(Wv(i, k), upL(&B(i, k), px, o), upR(&B(i, k), px, o)); }
                if (k + 1 <= nz - 1) { int o = ord_up_face(k + 1, nz); Fz1 = upwind(Wv(i, k + 1), upL(&B(i, k + 1), px, o), upR(&B(i, k + 1), px, o)); }
                /* ghost cells through the wall value (ValueBoundaryCondition) */
                double bdn = (k == 0) ? 2 * W->Tb[i] - B(i, 0) : B(i, k - 1);
                double bup = (k == nz - 1) ? 2 * P->b_top - B(i, nz - 1) : B(i, k + 1);
                double lap = (B(i + 1, k) - 2 * B(i, k) + B(i - 1, k)) * idx * idx + (bup - 2 * B(i, k) + bdn) * idz * idz;
                W->Gb[(size_t)k * nx + i] = -((Fx1 - Fx0) * idx + (Fz1 - Fz0) * idz) + P->kappa * lap;
            }
            /* ---------------- u at (x-face i, centre k) ---------------- */
            {
                /* Uu at centres i-1 and i:  centre c lies between faces c and c+1 */
                double ut0 = ce4(&U(i, k), 1), ut1 = ce4(&U(i + 1, k), 1);
                double F0 = upwind(ut0, up5L(&U(i, k), 1), up5R(&U(i, k), 1));            /* centre i-1 */
                double F1 = upwind(ut1, up5L(&U(i + 1, k), 1), up5R(&U(i + 1, k), 1));    /* centre i   */
                /* Wu at (x-face i, z-face k) and (x-face i, z-face k+1) */
                double H0 = 0.0, H1 = 0.0;
                if (k >= 1) { int o = ord_up_face(k, nz); double wt = ce4(&Wv(i, k), 1); H0 = upwind(wt, upL(&U(i, k), px, o), upR(&U(i, k), px, o)); }
                if (k + 1 <= nz - 1) { int o = ord_up_face(k + 1, nz); double wt = ce4(&Wv(i, k + 1), 1); H1 = upwind(wt, upL(&U(i, k + 1), px, o), upR(&U(i, k + 1), px, o)); }
                double udn = (k == 0) ? -U(i, 0) : U(i, k - 1);             /* no-slip: ghost = -interior */
                double uup = (k == nz - 1) ? -U(i, nz - 1) : U(i, k + 1);
                double lap = (U(i + 1, k) - 2 * U(i, k) + U(i - 1, k)) * idx * idx + (uup - 2 * U(i, k) + udn) * idz * idz;
                double g = -((F1 - F0) * idx + (H1 - H0) * idz) + P->nu * lap;
                if (P->split_phy) g -= (PH(i, k) - PH(i - 1, k)) * idx;
                W->Gu[(size_t)k * nx + i] = g;
            }
            /* ---------------- w at (centre i, z-face k), interior faces only ---------------- */
            if (k >= 1) {
                /* Uw at (x-face i, z-face k) and (x-face i+1, z-face k) */
                int oc = ord_ce_face(k, nz);
                double ut0 = cen(&U(i, k), px, oc), ut1 = cen(&U(i + 1, k), px, oc);
                double F0 = upwind(ut0, up5L(&Wv(i, k), 1), up5R(&Wv(i, k), 1));
                double F1 = upwind(ut1, up5L(&Wv(i + 1, k), 1), up5R(&Wv(i + 1, k), 1));
                /* Ww at centres k-1 and k: centre c lies between faces c and c+1 */
                int o0 = ord_up_cen(k - 1, nz), c0 = ord_ce_cen(k - 1, nz);
                int o1 = ord_up_cen(k, nz), c1 = ord_ce_cen(k, nz);
                double H0 = upwind(cen(&Wv(i, k), px, c0), upL(&Wv(i, k), px, o0), upR(&Wv(i, k), px, o0));
                double H1 = upwind(cen(&Wv(i, k + 1), px, c1), upL(&Wv(i, k + 1), px, o1), upR(&Wv(i, k + 1), px, o1));
                double lap = (Wv(i + 1, k) - 2 * Wv(i, k) + Wv(i - 1, k)) * idx * idx + (Wv(i, k + 1) - 2 * Wv(i, k) + Wv(i, k - 1)) * idz * idz;
                double g = -((F1 - F0) * idx + (H1 - H0) * idz) + P->nu * lap;
                if (!P->split_phy) g += 0.5 * (B(i, k - 1) + B(i, k));
                W->Gw[(size_t)k * nx + i] = g;
            }
        }
    }
    for (int i = 0; i < nx; ++i) { W->Gw[i] = 0.0; W->Gw[(size_t)nz * nx + i] = 0.0; }
#undef B
#undef U
#undef Wv
#undef PH
}

/* ------------------------------------------------------------------------------------- */
/* Poisson:  lap(p) = rhs, periodic x, Neumann z, zero-mean gauge (SURVEY 8a "Poisson")     */
/* ------------------------------------------------------------------------------------- */
static void dft_rows(const work_t *W, const double *in_re, const double *in_im, double *out_re, double *out_im, int sign)
{
    /* plain O(n^2) DFT along x for each row; exact restatement, used by both modes */
    const int nx = W->nx, nz = W->nz;
    for (int k = 0; k < nz; ++k) {
        const double *ar = in_re + (size_t)k * nx, *ai = in_im ? in_im + (size_t)k * nx : NULL;
        for (int m = 0; m < nx; ++m) {
            double sr = 0.0, si = 0.0;
            for (int i = 0; i < nx; ++i) {
                int j = (int)(((long)m * i) % nx);
                double c = W->cosx[j], s = sign * W->sinx[j];
                double xr = ar[i], xi = ai ? ai[i] : 0.0;
                sr += xr * c - xi * s;
                si += xr * s + xi * c;
            }
            out_re[(size_t)k * nx + m] = sr;
            out_im[(size_t)k * nx + m] = si;
        }
    }
}

/* recursive mixed-radix decimation-in-time FFT (any n; cost n*sum(prime factors)); twiddles from the
 * nx-point table: W_n^j = W_nx^(j*nx/n).  out[k + (n/p)q] = sum_r W_n^(r(k+(n/p)q)) Y_r[k].          */
static void fft_rec(const work_t *W, int n, int stride, const double *xr, const double *xi,
                    double *yr, double *yi, int sign, double *tmp)
{
    if (n == 1) { yr[0] = xr[0]; yi[0] = xi ? xi[0] : 0.0; return; }
    int p = 2;
    while (n % p) ++p;
    const int m = n / p, nx = W->nx, tw = nx / n;
    double *sr = tmp, *si = tmp + n;
    for (int r = 0; r < p; ++r)
        fft_rec(W, m, stride * p, xr + (size_t)r * stride, xi ? xi + (size_t)r * stride : NULL,
                sr + (size_t)r * m, si + (size_t)r * m, sign, tmp + 2 * n);
    if (p == 2) {
        for (int k = 0; k < m; ++k) {
            const double c = W->cosx[k * tw], s = sign * W->sinx[k * tw];
            const double tr = sr[m + k] * c - si[m + k] * s, ti = sr[m + k] * s + si[m + k] * c;
            yr[k] = sr[k] + tr; yi[k] = si[k] + ti;
            yr[k + m] = sr[k] - tr; yi[k + m] = si[k] - ti;
        }
        return;
    }
    if (p == 3) {
        const double h = sign * 0.86602540378443864676;   /* sin(2pi/3) */
        for (int k = 0; k < m; ++k) {
            const double c1 = W->cosx[k * tw], s1 = sign * W->sinx[k * tw];
            const double c2 = W->cosx[2 * k * tw], s2 = sign * W->sinx[2 * k * tw];
            const double ar = sr[m + k] * c1 - si[m + k] * s1, ai = sr[m + k] * s1 + si[m + k] * c1;
            const double br = sr[2 * m + k] * c2 - si[2 * m + k] * s2, bi = sr[2 * m + k] * s2 + si[2 * m + k] * c2;
            const double pr = ar + br, pi = ai + bi, mr = ar - br, mi = ai - bi;
            yr[k] = sr[k] + pr; yi[k] = si[k] + pi;
            const double xr = sr[k] - 0.5 * pr, xi = si[k] - 0.5 * pi;
            yr[k + m] = xr - h * mi; yi[k + m] = xi + h * mr;
            yr[k + 2 * m] = xr + h * mi; yi[k + 2 * m] = xi - h * mr;
        }
        return;
    }
    for (int q = 0; q < p; ++q)
        for (int k = 0; k < m; ++k) {
            const int kk = k + m * q;
            double ar = 0.0, ai = 0.0;
            for (int r = 0; r < p; ++r) {
                const int j = (int)(((long)r * kk * tw) % nx);
                const double c = W->cosx[j], s = sign * W->sinx[j];
                const double vr = sr[(size_t)r * m + k], vi = si[(size_t)r * m + k];
                ar += vr * c - vi * s;
                ai += vr * s + vi * c;
            }
            yr[kk] = ar; yi[kk] = ai;
        }
}

static void fft_rows(const work_t *W, const double *in_re, const double *in_im, double *out_re, double *out_im, int sign)
{
    const int nx = W->nx, nz = W->nz;
    double *tmp = (double *)malloc((size_t)nx * 8 * 8);
    for (int k = 0; k < nz; ++k)
        fft_rec(W, nx, 1, in_re + (size_t)k * nx, in_im ? in_im + (size_t)k * nx : NULL,
                out_re + (size_t)k * nx, out_im + (size_t)k * nx, sign, tmp);
    free(tmp);
}

static void poisson(const rbc2d_params *P, work_t *W)
{
    const int nx = P->nx, nz = P->nz;
    const double dx = P->lx / nx, dz = P->lz / nz;
    double *pre = (double *)malloc((size_t)nx * nz * 8), *pim = (double *)malloc((size_t)nx * nz * 8);
    if (P->poisson_mode == 0) fft_rows(W, W->rhs, NULL, W->fre, W->fim, -1);
    else dft_rows(W, W->rhs, NULL, W->fre, W->fim, -1);
    if (P->poisson_mode == 0) {
        /* per x-wavenumber: (p(k+1)-2p(k)+p(k-1))/dz^2 - lx p(k) = r(k), p(-1)=p(0), p(nz)=p(nz-1) */
        double *cp = (double *)malloc(nz * 8), *dr = (double *)malloc(nz * 8), *di = (double *)malloc(nz * 8);
        for (int m = 0; m < nx; ++m) {
            double sx = 2 * sin(M_PI * m / nx) / dx, lam = sx * sx * dz * dz;
            for (int k = 0; k < nz; ++k) {
                double diag = -(2.0 + lam);
                if (k == 0 || k == nz - 1) diag += 1.0;
                if (m == 0 && k == 0) diag -= 1.0;           /* pin the null space: p(0)=0, mean removed below */
                double fr = W->fre[(size_t)k * nx + m] * dz * dz, fi = W->fim[(size_t)k * nx + m] * dz * dz;
                if (k == 0) { cp[0] = 1.0 / diag; dr[0] = fr / diag; di[0] = fi / diag; }
                else {
                    double den = diag - cp[k - 1];
                    cp[k] = 1.0 / den; dr[k] = (fr - dr[k - 1]) / den; di[k] = (fi - di[k - 1]) / den;
                }
            }
            pre[(size_t)(nz - 1) * nx + m] = dr[nz - 1]; pim[(size_t)(nz - 1) * nx + m] = di[nz - 1];
            for (int k = nz - 2; k >= 0; --k) {
                pre[(size_t)k * nx + m] = dr[k] - cp[k] * pre[(size_t)(k + 1) * nx + m];
                pim[(size_t)k * nx + m] = di[k] - cp[k] * pim[(size_t)(k + 1) * nx + m];
            }
        }
        free(cp); free(dr); free(di);
        double mr = 0.0;                                     /* zero-mean gauge: remove kz=0 of m=0 */
        for (int k = 0; k < nz; ++k) mr += pre[(size_t)k * nx];
        mr /= nz;
        for (int k = 0; k < nz; ++k) { pre[(size_t)k * nx] -= mr; pim[(size_t)k * nx] = 0.0; }
    } else {
        /* literal reference formulation: DCT-II in z, divide by -(lx+lz), (0,0) mode := 0 */
        double *cr = (double *)malloc((size_t)nx * nz * 8), *ci = (double *)malloc((size_t)nx * nz * 8);
        for (int q = 0; q < nz; ++q)
            for (int m = 0; m < nx; ++m) {
                double sr = 0.0, si = 0.0;
                for (int k = 0; k < nz; ++k) {
                    double c = cos(M_PI * (k + 0.5) * q / nz);
                    sr += W->fre[(size_t)k * nx + m] * c; si += W->fim[(size_t)k * nx + m] * c;
                }
                double sx = 2 * sin(M_PI * m / nx) / dx, sz = 2 * sin(M_PI * q / (2.0 * nz)) / dz;
                double lam = sx * sx + sz * sz;
                if (m == 0 && q == 0) { sr = 0.0; si = 0.0; } else { sr = -sr / lam; si = -si / lam; }
                cr[(size_t)q * nx + m] = sr; ci[(size_t)q * nx + m] = si;
            }
        for (int k = 0; k < nz; ++k)
            for (int m = 0; m < nx; ++m) {
                double sr = cr[m], si = ci[m];
                for (int q = 1; q < nz; ++q) {
                    double c = 2 * cos(M_PI * (k + 0.5) * q / nz);
                    sr += cr[(size_t)q * nx + m] * c; si += ci[(size_t)q * nx + m] * c;
                }
                pre[(size_t)k * nx + m] = sr / nz; pim[(size_t)k * nx + m] = si / nz;
            }
        free(cr); free(ci);
    }
    if (P->poisson_mode == 0) fft_rows(W, pre, pim, W->fre, W->fim, +1);
    else dft_rows(W, pre, pim, W->fre, W->fim, +1);
    for (size_t n = 0; n < (size_t)nx * nz; ++n) W->p[n] = W->fre[n] / nx;
    free(pre); free(pim);
}

/* divergence -> rhs/dtau -> solve -> correct u,w (interior faces) */
static void project(const rbc2d_params *P, work_t *W, double *u, double *w, double dtau)
{
    const int nx = P->nx, nz = P->nz;
    const double dx = P->lx / nx, dz = P->lz / nz;
    for (int k = 0; k < nz; ++k)
        for (int i = 0; i < nx; ++i) {
            int ip = (i + 1) % nx;
            double d = (u[(size_t)k * nx + ip] - u[(size_t)k * nx + i]) / dx + (w[(size_t)(k + 1) * nx + i] - w[(size_t)k * nx + i]) / dz;
            W->rhs[(size_t)k * nx + i] = d / dtau;
        }
    poisson(P, W);
    for (int k = 0; k < nz; ++k)
        for (int i = 0; i < nx; ++i) {
            int im = (i + nx - 1) % nx;
            u[(size_t)k * nx + i] -= dtau * (W->p[(size_t)k * nx + i] - W->p[(size_t)k * nx + im]) / dx;
            if (k >= 1) w[(size_t)k * nx + i] -= dtau * (W->p[(size_t)k * nx + i] - W->p[(size_t)(k - 1) * nx + i]) / dz;
        }
}

/* ------------------------------------------------------------------------------------- */
/* RK3 time step (Oceananigans RungeKutta3TimeStepper; SURVEY 8a "Time stepping")           */
/* ------------------------------------------------------------------------------------- */
static void rk3_step(const rbc2d_params *P, work_t *W, double *b, double *u, double *w, double dt)
{
    static const double gam[3] = {8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0};
    static const double zet[3] = {0.0, -17.0 / 60.0, -5.0 / 12.0};
    const int nx = P->nx, nz = P->nz;
    const size_t n = (size_t)nx * nz;
    for (int s = 0; s < 3; ++s) {
        tendencies(P, W, b, u, w);
        const double g = gam[s], z = zet[s];
        for (size_t q = 0; q < n; ++q) {
            b[q] += dt * (g * W->Gb[q] + z * W->Gb0[q]);
            u[q] += dt * (g * W->Gu[q] + z * W->Gu0[q]);
        }
        for (size_t q = (size_t)nx; q < n; ++q) w[q] += dt * (g * W->Gw[q] + z * W->Gw0[q]);
        project(P, W, u, w, (g + z) * dt);
        double *t;
        t = W->Gb0; W->Gb0 = W->Gb; W->Gb = t;
        t = W->Gu0; W->Gu0 = W->Gu; W->Gu = t;
        t = W->Gw0; W->Gw0 = W->Gw; W->Gw = t;
    }
}

/* ------------------------------------------------------------------------------------- */
/* public API                                                                               */
/* ------------------------------------------------------------------------------------- */

/* One action step: step_simulation (rbc_sim2D_api.jl:75-97).  The substep schedule
 * (run!'s  dt' = min(dt_solver, stop_time - t)) is supplied by the caller as dts[nsub].
 * Optional outputs phy/pnhs (nz*nx) receive pHY' (split mode; recomputed from the final b as
 * update_state! does after the last stage) and the last stage's non-hydrostatic pressure.
 * Returns 0 if the final b,u,w (w faces 0..nz-1) are NaN-free, 1 otherwise (rbc_sim2D.jl:223-228). */
int rbc2d_oracle_step(const rbc2d_params *P, double *b, double *u, double *w, const double *action,
                      int nsub, const double *dts, double *phy, double *pnhs)
{
    work_t *W = work_new(P->nx, P->nz);
    rbc2d_oracle_heater_profile(P, action, W->Tb);
    for (int s = 0; s < nsub; ++s) rk3_step(P, W, b, u, w, dts[s]);
    const int nx = P->nx, nz = P->nz;
    if (pnhs) memcpy(pnhs, W->p, (size_t)nx * nz * 8);
    if (phy) {
        rbc2d_params Q = *P; Q.split_phy = 1;
        tendencies(&Q, W, b, u, w);
        for (int k = 0; k < nz; ++k)
            for (int i = 0; i < nx; ++i) phy[(size_t)k * nx + i] = W->phy[(size_t)k * W->px + HX + i];
    }
    int bad = 0;
    for (size_t q = 0; q < (size_t)nx * nz; ++q) if (isnan(b[q]) || isnan(u[q]) || isnan(w[q])) { bad = 1; break; }
    work_free(W);
    return bad;
}

/* Batched driver used only by bench.py's CPU baseline: B independent envs, OpenMP over envs. */
int rbc2d_oracle_step_batch(const rbc2d_params *P, int B, double *b, double *u, double *w, const double *actions,
                            int nsub, const double *dts, int nthreads)
{
    int bad = 0;
    const size_t n = (size_t)P->nx * P->nz, nwv = (size_t)P->nx * (P->nz + 1);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
    for (int e = 0; e < B; ++e)
        bad |= rbc2d_oracle_step(P, b + e * n, u + e * n, w + e * nwv, actions + (size_t)e * P->heaters, nsub, dts, NULL, NULL);
    return bad;
}

int rbc2d_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Tendencies of a given state (for the fixture residual tests). Outputs unpadded Gb,Gu (nz*nx), Gw ((nz+1)*nx). */
void rbc2d_oracle_tendencies(const rbc2d_params *P, const double *b, const double *u, const double *w, const double *action,
                             double *Gb, double *Gu, double *Gw)
{
    work_t *W = work_new(P->nx, P->nz);
    rbc2d_oracle_heater_profile(P, action, W->Tb);
    tendencies(P, W, b, u, w);
    memcpy(Gb, W->Gb, (size_t)P->nx * P->nz * 8);
    memcpy(Gu, W->Gu, (size_t)P->nx * P->nz * 8);
    memcpy(Gw, W->Gw, (size_t)P->nx * (P->nz + 1) * 8);
    work_free(W);
}

/* Projection alone, as Oceananigans' set! applies to a freshly initialised model (dt = 1). */
void rbc2d_oracle_project(const rbc2d_params *P, double *u, double *w, double *p_out)
{
    work_t *W = work_new(P->nx, P->nz);
    project(P, W, u, w, 1.0);
    if (p_out) memcpy(p_out, W->p, (size_t)P->nx * P->nz * 8);
    work_free(W);
}

/* get_nusselt (rbc_sim2D_api.jl:142-163) on a (T, w) pair laid out [nz][nx] (Python layout):
 * q1 = mean(T.*w); Tx = mean over x; q2 = kappa*mean(array_gradient(Tx)); Nu = (q1-q2)/(kappa*db/H) */
double rbc2d_oracle_nusselt(const double *T, const double *wv, int nz, int nx, double kappa, double db, double H)
{
    double q1 = 0.0;
    double *Tx = (double *)malloc(nz * 8);
    for (int k = 0; k < nz; ++k) {
        double s = 0.0;
        for (int i = 0; i < nx; ++i) { q1 += T[(size_t)k * nx + i] * wv[(size_t)k * nx + i]; s += T[(size_t)k * nx + i]; }
        Tx[k] = s / nx;
    }
    q1 /= (double)nx * nz;
    double g = 0.0;                         /* array_gradient, rbc_sim2D.jl:206-220 (unit spacing) */
    for (int k = 0; k < nz; ++k) {
        if (k == 0) g += Tx[1] - Tx[0];
        else if (k == nz - 1) g += Tx[k] - Tx[k - 1];
        else g += (Tx[k + 1] - Tx[k - 1]) / 2;
    }
    g /= nz;
    free(Tx);
    return (q1 - kappa * g) / (kappa * db / H);
}
