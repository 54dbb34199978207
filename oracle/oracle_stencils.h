/*
 * oracle_stencils.h — reconstruction stencils and wall-order rules shared by the 2D and 3D CPU oracles
 * (TEST INFRASTRUCTURE ONLY; see rbc2d_oracle.c).  Scheme: SURVEY.md section 8a "Advection".
 */
#ifndef ORACLE_STENCILS_H
#define ORACLE_STENCILS_H
#include <math.h>

/* ------------------------------------------------------------------------------------- */
/* reconstruction stencils (SURVEY 8a "Advection")                                         */
/* face between cells j-1 | j ; c points at cell j, stride s                               */
/* ------------------------------------------------------------------------------------- */
static inline double up5L(const double *c, long s) { return (2*c[-3*s] - 13*c[-2*s] + 47*c[-s] + 27*c[0] - 3*c[s]) / 60.0; }
static inline double up5R(const double *c, long s) { return (-3*c[-2*s] + 27*c[-s] + 47*c[0] - 13*c[s] + 2*c[2*s]) / 60.0; }
static inline double up3L(const double *c, long s) { return (-c[-2*s] + 5*c[-s] + 2*c[0]) / 6.0; }
static inline double up3R(const double *c, long s) { return (2*c[-s] + 5*c[0] - c[s]) / 6.0; }
static inline double up1L(const double *c, long s) { return c[-s]; }
static inline double up1R(const double *c, long s) { (void)s; return c[0]; }
static inline double ce4(const double *c, long s)  { return (-c[-2*s] + 7*c[-s] + 7*c[0] - c[s]) / 12.0; }
static inline double ce2(const double *c, long s)  { return (c[-s] + c[0]) / 2.0; }

static inline double upwind(double ut, double pL, double pR)
{
    return 0.5 * ((ut + fabs(ut)) * pL + (ut - fabs(ut)) * pR);
}

/* order of the biased stencil at z-face kf (centres -> face); faces 1..nz-1 interior */
static inline int ord_up_face(int kf, int nz)  { return (kf >= 3 && kf <= nz - 3) ? 5 : (kf == 2 || kf == nz - 2) ? 3 : 1; }
static inline int ord_ce_face(int kf, int nz)  { return (kf >= 2 && kf <= nz - 2) ? 4 : 2; }
/* order at centre kc (z-faces -> centre) */
static inline int ord_up_cen(int kc, int nz)   { return (kc >= 2 && kc <= nz - 3) ? 5 : (kc == 1 || kc == nz - 2) ? 3 : 1; }
static inline int ord_ce_cen(int kc, int nz)   { return (kc >= 1 && kc <= nz - 2) ? 4 : 2; }

static inline double upL(const double *c, long s, int ord) { return ord == 5 ? up5L(c, s) : ord == 3 ? up3L(c, s) : up1L(c, s); }
static inline double upR(const double *c, long s, int ord) { return ord == 5 ? up5R(c, s) : ord == 3 ? up3R(c, s) : up1R(c, s); }
static inline double cen(const double *c, long s, int ord) { return ord == 4 ? ce4(c, s) : ce2(c, s); }


#endif
