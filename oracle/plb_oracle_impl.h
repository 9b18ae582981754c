/* oracle/plb_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (see plb_oracle.h).
 *
 * Included twice by plb_oracle.c, with  REAL = double, SFX(x) = x##_f64  and
 * REAL = float, SFX(x) = x##_f32.  Every function restates one function of the
 * reference and cites it (paths relative to /root/reference/src).  Expression
 * shapes -- association order, where a literal is (T)-cast and where it is a bare
 * double -- are kept exactly as in the reference, because the parity contract is
 * bit-exactness of populations (compiled with -ffp-contract=off; x86-64 SSE2 has no
 * implicit FMA and FLT_EVAL_METHOD == 0).
 */

typedef struct { int type; REAL omega, cSmago, preFactor, rho; } SFX(Dyn);
typedef struct { int kind, dir, ori; REAL u[3]; REAL rhoBar; int autoPrep; /* automaticPrepareCollision, core/dynamics.hh:470-475 */ } SFX(Bc);

typedef struct {
    int q, nx, ny, nz, per[3];
    const int (*c)[3];
    REAL t[27];
    REAL cs2, invCs2;
    REAL *f, *g;                 /* AoS [x][y][z][q], as BlockLattice3D (atomicBlock/blockLattice3D.hh:334-350) */
    uint8_t* dynIdx;             /* per cell: index into dyns[] (the cell's bulk / base dynamics) */
    int32_t* bcIdx;              /* per cell: index into bcs[] or -1 (the composite boundary dynamics on top) */
    SFX(Dyn) dyns[MAX_DYN]; int nDyn;
    SFX(Bc)* bcs; int nBc, capBc;
    Proc* procs; int nProc, capProc;
    int statBox[6];
    REAL* ext;                   /* external rhoBar / j fields [x][y][z][rhoBar,jx,jy,jz] (ScalarField3D + TensorField3D<3>) or NULL */
    double tmpSumRhoBar, tmpSumUSqr, tmpMaxUSqr; uint64_t tmpN;   /* running (core/blockStatistics.cpp:205-229) */
    double avRhoBar, avUSqr, maxUSqr; uint64_t numCells;          /* public, after evaluate() (:77-120) */
} SFX(Lat);

#define CELL(L, x, y, z) ((size_t)(((size_t)(x) * (L)->ny + (y)) * (L)->nz + (z)))

/* ---------------------------------------------------------------- descriptor ---- */

/* latticeBoltzmann/nearestNeighborLattices3D.hh:176-196 (D3Q19 t), :263-277 (D3Q27 t) */
static void SFX(init_descriptor)(SFX(Lat)* L) {
    int i;
    if (L->q == 19) {
        L->c = c19;
        L->t[0] = (REAL)1/(REAL)3;
        for (i = 1; i < 19; ++i) L->t[i] = (c19n[i] == 1) ? (REAL)1/(REAL)18 : (REAL)1/(REAL)36;
    } else {
        L->c = c27;
        L->t[0] = (REAL)8/(REAL)27;
        for (i = 1; i < 27; ++i)
            L->t[i] = (c27n[i] == 1) ? (REAL)2/(REAL)27 : (c27n[i] == 2) ? (REAL)1/(REAL)54 : (REAL)1/(REAL)216;
    }
    L->cs2 = (REAL)1 / (REAL)3;
    L->invCs2 = (REAL)3;
}

/* latticeBoltzmann/roundOffPolicy.h:74-90 (DefaultRoundOffPolicy) */
#define RHOBAR(rho)     ((rho) - (REAL)1)
#define FULLRHO(rhoBar) ((rhoBar) + (REAL)1)
#define INVRHO(rhoBar)  ((REAL)1 / ((rhoBar) + (REAL)1))

/* ------------------------------------------------------------------- moments ---- */

/* latticeBoltzmann/momentTemplates3D.h:351-373 (D3Q19), :43-72 (D3Q27) */
static void SFX(partial_rho)(int q, const REAL* f, REAL* xm, REAL* x0, REAL* xp,
                             REAL* ym, REAL* yp, REAL* zm, REAL* zp)
{
    if (q == 19) {
        *xm = f[1] + f[4] + f[5] + f[6] + f[7];
        *xp = f[10] + f[13] + f[14] + f[15] + f[16];
        *ym = f[2] + f[4] + f[8] + f[9] + f[14];
        *yp = f[5] + f[11] + f[13] + f[17] + f[18];
        *zm = f[3] + f[6] + f[8] + f[16] + f[18];
        *zp = f[7] + f[9] + f[12] + f[15] + f[17];
        if (x0) *x0 = f[0] + f[2] + f[3] + f[8] +
                      f[9] + f[11] + f[12] + f[17] + f[18];
    } else {
        *xm = f[1] + f[4] + f[5] + f[6] + f[7] +
              f[10] + f[11] + f[12] + f[13];
        *xp = f[14] + f[17] + f[18] + f[19] + f[20] +
              f[23] + f[24] + f[25] + f[26];
        *ym = f[2] + f[4] + f[8] + f[9] + f[10] +
              f[11] + f[18] + f[25] + f[26];
        *yp = f[5] + f[12] + f[13] + f[15] + f[17] +
              f[21] + f[22] + f[23] + f[24];
        *zm = f[3] + f[6] + f[8] + f[10] + f[12] +
              f[20] + f[22] + f[24] + f[26];
        *zp = f[7] + f[9] + f[11] + f[13] + f[16] +
              f[19] + f[21] + f[23] + f[25];
        if (x0) *x0 = f[0] + f[2] + f[3] + f[8] +
                      f[9] + f[15] + f[16] + f[21] + f[22];
    }
}

/* momentTemplates3D.h:375-382 (D3Q19 get_rhoBar), :74-83 (D3Q27): plain left-to-right sum */
static REAL SFX(get_rhoBar)(int q, const REAL* f) {
    REAL rhoBar = f[0] + f[1];
    int i;
    for (i = 2; i < q; ++i) rhoBar = rhoBar + f[i];
    return rhoBar;
}

/* momentTemplates3D.h:384-402 (get_j) */
static void SFX(get_j)(int q, const REAL* f, REAL j[3]) {
    REAL xm, xp, ym, yp, zm, zp;
    SFX(partial_rho)(q, f, &xm, 0, &xp, &ym, &yp, &zm, &zp);
    j[0] = (xp - xm); j[1] = (yp - ym); j[2] = (zp - zm);
}

/* momentTemplates3D.h:411-423 (D3Q19 get_rhoBar_j), :118-132 (D3Q27) */
static void SFX(get_rhoBar_j)(int q, const REAL* f, REAL* rhoBar, REAL j[3]) {
    REAL xm, x0, xp, ym, yp, zm, zp;
    SFX(partial_rho)(q, f, &xm, &x0, &xp, &ym, &yp, &zm, &zp);
    *rhoBar = xm + x0 + xp;
    j[0] = (xp - xm); j[1] = (yp - ym); j[2] = (zp - zm);
}

/* momentTemplates3D.h:517-539 (D3Q19 compute_rhoBar_j_PiNeq), :239-267 (D3Q27).
 * Tensor index order xx,xy,xz,yy,yz,zz (geometricOperationTemplates.h:246-248). */
static void SFX(rhoBar_j_PiNeq)(int q, const REAL* f, REAL cs2, REAL* rhoBar, REAL j[3], REAL PiNeq[6]) {
    REAL xm, x0, xp, ym, yp, zm, zp, invRho;
    SFX(partial_rho)(q, f, &xm, &x0, &xp, &ym, &yp, &zm, &zp);
    *rhoBar = xm + x0 + xp;
    j[0] = (xp - xm); j[1] = (yp - ym); j[2] = (zp - zm);
    invRho = INVRHO(*rhoBar);
    PiNeq[0] = xp+xm - cs2*(*rhoBar) - invRho*j[0]*j[0];
    PiNeq[3] = yp+ym - cs2*(*rhoBar) - invRho*j[1]*j[1];
    PiNeq[5] = zp+zm - cs2*(*rhoBar) - invRho*j[2]*j[2];
    if (q == 19) {
        PiNeq[1] = f[4] - f[5] + f[13] - f[14] - invRho*j[0]*j[1];
        PiNeq[2] = f[6] - f[7] + f[15] - f[16] - invRho*j[0]*j[2];
        PiNeq[4] = f[8] - f[9] + f[17] - f[18] - invRho*j[1]*j[2];
    } else {
        PiNeq[1] = f[4] - f[5] + f[10] + f[11] - f[12] - f[13]
                   + f[17] - f[18] + f[23] + f[24] - f[25] - f[26]
                   - invRho*j[0]*j[1];
        PiNeq[2] = f[6] - f[7] + f[10] - f[11] + f[12] - f[13]
                   + f[19] - f[20] + f[23] - f[24] + f[25] - f[26]
                   - invRho*j[0]*j[2];
        PiNeq[4] = f[8] - f[9] + f[10] - f[11] - f[12] + f[13]
                   + f[21] - f[22] + f[23] - f[24] - f[25] + f[26]
                   -invRho*j[1]*j[2];
    }
}

/* momentTemplates<T,Descriptor>::compute_PiNeq(cell, rhoBar, j, PiNeq) latticeBoltzmann/momentTemplates.h:118-123 ->
 * momentTemplatesImpl::compute_PiNeq(f, rhoBar, j, PiNeq, invRho) momentTemplates3D.h:482-497 (D3Q19), :191-212 (D3Q27):
 * the off-equilibrium stress from the populations and EXTERNALLY supplied moments. */
static void SFX(compute_PiNeq)(int q, const REAL* f, REAL cs2, REAL rhoBar, const REAL j[3], REAL PiNeq[6]) {
    REAL xm, xp, ym, yp, zm, zp, invRho = INVRHO(rhoBar);
    SFX(partial_rho)(q, f, &xm, 0, &xp, &ym, &yp, &zm, &zp);
    PiNeq[0] = xp+xm - cs2*rhoBar - invRho*j[0]*j[0];
    PiNeq[3] = yp+ym - cs2*rhoBar - invRho*j[1]*j[1];
    PiNeq[5] = zp+zm - cs2*rhoBar - invRho*j[2]*j[2];
    if (q == 19) {
        PiNeq[1] = f[4] - f[5] + f[13] - f[14] - invRho*j[0]*j[1];
        PiNeq[2] = f[6] - f[7] + f[15] - f[16] - invRho*j[0]*j[2];
        PiNeq[4] = f[8] - f[9] + f[17] - f[18] - invRho*j[1]*j[2];
    } else {
        PiNeq[1] = f[4] - f[5] + f[10] + f[11] - f[12] - f[13]
                   + f[17] - f[18] + f[23] + f[24] - f[25] - f[26]
                   - invRho*j[0]*j[1];
        PiNeq[2] = f[6] - f[7] + f[10] - f[11] + f[12] - f[13]
                   + f[19] - f[20] + f[23] - f[24] + f[25] - f[26]
                   - invRho*j[0]*j[2];
        PiNeq[4] = f[8] - f[9] + f[10] - f[11] - f[12] + f[13]
                   + f[21] - f[22] + f[23] - f[24] - f[25] + f[26]
                   -invRho*j[1]*j[2];
    }
}

/* ---------------------------------------------------------------- equilibria ---- */

/* dynamicsTemplates3D.h:1994-1997 (D3Q19: literals are (T)-cast) and :167-170 (D3Q27: bare
 * double literals, so a float lattice evaluates this in double -- SURVEY.md section 7). */
static REAL SFX(bgk_ma2_equilibrium)(const SFX(Lat)* L, int i, REAL rhoBar, REAL invRho, const REAL j[3], REAL jSqr) {
    REAL c_j = L->c[i][0]*j[0] + L->c[i][1]*j[1] + L->c[i][2]*j[2];
    if (L->q == 19)
        return L->t[i] * ( rhoBar + (REAL)3.*c_j + invRho*((REAL)4.5*c_j*c_j - (REAL)1.5*jSqr) );
    return L->t[i] * ( rhoBar + 3.*c_j + invRho*(4.5*c_j*c_j - 1.5*jSqr) );
}

/* Dynamics::computeEquilibrium(iPop, rhoBar, j, jSqr): BGK isoThermalDynamics.hh:341-346,
 * RLB :1619-1624, TRT trtDynamics.hh:138-143, Smagorinsky smagorinskyDynamics.hh:322-328
 * (all the same); BounceBack / NoDynamics return 0 (core/dynamics.hh:800-805, :980-985). */
static REAL SFX(computeEquilibrium)(const SFX(Lat)* L, const SFX(Dyn)* d, int i, REAL rhoBar, const REAL j[3], REAL jSqr) {
    REAL invRho;
    if (d->type == PLBO_DYN_BOUNCEBACK || d->type == PLBO_DYN_NODYNAMICS) return (REAL)0;
    /* IncBGKdynamics::computeEquilibrium isoThermalDynamics.hh:1428-1435: the 1/rho of the O(Ma^2) term is 1/rho0 = 1 */
    invRho = d->type == PLBO_DYN_INCBGK ? (REAL)1 : INVRHO(rhoBar);
    return SFX(bgk_ma2_equilibrium)(L, i, rhoBar, invRho, j, jSqr);
}

/* dynamicsTemplates3D.h:1999-2076 (D3Q19 bgk_ma2_equilibria), :264-365 (D3Q27) */
static void SFX(bgk_ma2_equilibria)(const SFX(Lat)* L, REAL rhoBar, REAL invRho, const REAL j[3], REAL jSqr, REAL* eq) {
    const int h = L->q / 2;
    REAL t0 = L->t[0], t1 = L->t[1], t4 = L->t[4];
    REAL kx = (REAL)3 * j[0], ky = (REAL)3 * j[1], kz = (REAL)3 * j[2];
    REAL kxSqr_ = invRho / (REAL)2 * kx*kx;
    REAL kySqr_ = invRho / (REAL)2 * ky*ky;
    REAL kzSqr_ = invRho / (REAL)2 * kz*kz;
    REAL kxky_  = invRho * kx*ky;
    REAL kxkz_  = invRho * kx*kz;
    REAL kykz_  = invRho * ky*kz;
    REAL C1 = rhoBar + invRho*(REAL)3*jSqr;
    REAL C2, C3;
    C3 = -kxSqr_ - kySqr_ - kzSqr_;           eq[0] = t0 * (C1+C3);
    C2 = -kx;       C3 = -kySqr_ - kzSqr_;    eq[1] = t1 * (C1+C2+C3); eq[1+h] = t1 * (C1-C2+C3);
    C2 = -ky;       C3 = -kxSqr_ - kzSqr_;    eq[2] = t1 * (C1+C2+C3); eq[2+h] = t1 * (C1-C2+C3);
    C2 = -kz;       C3 = -kxSqr_ - kySqr_;    eq[3] = t1 * (C1+C2+C3); eq[3+h] = t1 * (C1-C2+C3);
    C2 = -kx - ky;  C3 = kxky_ - kzSqr_;      eq[4] = t4 * (C1+C2+C3); eq[4+h] = t4 * (C1-C2+C3);
    C2 = -kx + ky;  C3 = -kxky_ - kzSqr_;     eq[5] = t4 * (C1+C2+C3); eq[5+h] = t4 * (C1-C2+C3);
    C2 = -kx - kz;  C3 = kxkz_ - kySqr_;      eq[6] = t4 * (C1+C2+C3); eq[6+h] = t4 * (C1-C2+C3);
    C2 = -kx + kz;  C3 = -kxkz_ - kySqr_;     eq[7] = t4 * (C1+C2+C3); eq[7+h] = t4 * (C1-C2+C3);
    C2 = -ky - kz;  C3 = kykz_ - kxSqr_;      eq[8] = t4 * (C1+C2+C3); eq[8+h] = t4 * (C1-C2+C3);
    C2 = -ky + kz;  C3 = -kykz_ - kxSqr_;     eq[9] = t4 * (C1+C2+C3); eq[9+h] = t4 * (C1-C2+C3);
    if (L->q == 27) {
        REAL t10 = L->t[10];
        C2 = -kx -ky -kz;  C3 = kxky_ + kxkz_ + kykz_;   eq[10] = t10 * (C1+C2+C3); eq[23] = t10 * (C1-C2+C3);
        C2 = -kx -ky +kz;  C3 = kxky_ - kxkz_ - kykz_;   eq[11] = t10 * (C1+C2+C3); eq[24] = t10 * (C1-C2+C3);
        C2 = -kx +ky -kz;  C3 = -kxky_ + kxkz_ - kykz_;  eq[12] = t10 * (C1+C2+C3); eq[25] = t10 * (C1-C2+C3);
        C2 = -kx +ky +kz;  C3 = -kxky_ - kxkz_ + kykz_;  eq[13] = t10 * (C1+C2+C3); eq[26] = t10 * (C1-C2+C3);
    }
}

/* ---------------------------------------------------------------- collisions ---- */

/* D3Q19: dynamicsTemplates3D.h:2289-2368 (bgk_ma2_collision_base) via :2476-2478.
 * D3Q27: generic loop :1624-1634.  Returns uSqr. */
static REAL SFX(bgk_ma2_collision)(const SFX(Lat)* L, REAL* f, REAL rhoBar, const REAL j[3], REAL omega) {
    REAL invRho = INVRHO(rhoBar);
    if (L->q == 19) {
        REAL one_m_omega = (REAL)1 - omega;
        REAL t0_omega = L->t[0] * omega;
        REAL t1_omega = L->t[1] * omega;
        REAL t4_omega = L->t[4] * omega;
        REAL jSqr   = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
        REAL kx     = (REAL)3 * j[0];
        REAL ky     = (REAL)3 * j[1];
        REAL kz     = (REAL)3 * j[2];
        REAL kxSqr_ = invRho / (REAL)2 * kx*kx;
        REAL kySqr_ = invRho / (REAL)2 * ky*ky;
        REAL kzSqr_ = invRho / (REAL)2 * kz*kz;
        REAL kxky_  = invRho * kx*ky;
        REAL kxkz_  = invRho * kx*kz;
        REAL kykz_  = invRho * ky*kz;
        REAL C1 = rhoBar + invRho*(REAL)3*jSqr;
        REAL C2, C3;
        C3 = -kxSqr_ - kySqr_ - kzSqr_;
        f[0] *= one_m_omega; f[0] += t0_omega * (C1+C3);
        C2 = -kx;  C3 = -kySqr_ - kzSqr_;
        f[1]  *= one_m_omega; f[1]  += t1_omega * (C1+C2+C3);
        f[10] *= one_m_omega; f[10] += t1_omega * (C1-C2+C3);
        C2 = -ky;  C3 = -kxSqr_ - kzSqr_;
        f[2]  *= one_m_omega; f[2]  += t1_omega * (C1+C2+C3);
        f[11] *= one_m_omega; f[11] += t1_omega * (C1-C2+C3);
        C2 = -kz;  C3 = -kxSqr_ - kySqr_;
        f[3]  *= one_m_omega; f[3]  += t1_omega * (C1+C2+C3);
        f[12] *= one_m_omega; f[12] += t1_omega * (C1-C2+C3);
        C2 = -kx - ky;  C3 = kxky_ - kzSqr_;
        f[4]  *= one_m_omega; f[4]  += t4_omega * (C1+C2+C3);
        f[13] *= one_m_omega; f[13] += t4_omega * (C1-C2+C3);
        C2 = -kx + ky;  C3 = -kxky_ - kzSqr_;
        f[5]  *= one_m_omega; f[5]  += t4_omega * (C1+C2+C3);
        f[14] *= one_m_omega; f[14] += t4_omega * (C1-C2+C3);
        C2 = -kx - kz;  C3 = kxkz_ - kySqr_;
        f[6]  *= one_m_omega; f[6]  += t4_omega * (C1+C2+C3);
        f[15] *= one_m_omega; f[15] += t4_omega * (C1-C2+C3);
        C2 = -kx + kz;  C3 = -kxkz_ - kySqr_;
        f[7]  *= one_m_omega; f[7]  += t4_omega * (C1+C2+C3);
        f[16] *= one_m_omega; f[16] += t4_omega * (C1-C2+C3);
        C2 = -ky - kz;  C3 = kykz_ - kxSqr_;
        f[8]  *= one_m_omega; f[8]  += t4_omega * (C1+C2+C3);
        f[17] *= one_m_omega; f[17] += t4_omega * (C1-C2+C3);
        C2 = -ky + kz;  C3 = -kykz_ - kxSqr_;
        f[9]  *= one_m_omega; f[9]  += t4_omega * (C1+C2+C3);
        f[18] *= one_m_omega; f[18] += t4_omega * (C1-C2+C3);
        return invRho*invRho*jSqr;
    } else {
        const REAL jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
        int i;
        for (i = 0; i < 27; ++i) {
            f[i] *= (REAL)1-omega;
            f[i] += omega * SFX(bgk_ma2_equilibrium)(L, i, rhoBar, invRho, j, jSqr);
        }
        return jSqr*invRho*invRho;
    }
}

/* bgk_ma2_collision_base(f, rhoBar, j, omega, invRho): dynamicsTemplates3D.h:2289-2368 (D3Q19), :467-571 (D3Q27);
 * what bgk_inc_collision (:2480-2482, :1680-1682) calls with invRho0.  Pairs (i, i+q/2) in index order. */
static REAL SFX(bgk_ma2_collision_base)(const SFX(Lat)* L, REAL* f, REAL rhoBar, const REAL j[3], REAL omega, REAL invRho) {
    const int h = L->q / 2;
    REAL one_m_omega = (REAL)1 - omega;
    REAL t0_omega = L->t[0] * omega;
    REAL t1_omega = L->t[1] * omega;
    REAL t4_omega = L->t[4] * omega;
    REAL jSqr   = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
    REAL kx     = (REAL)3 * j[0];
    REAL ky     = (REAL)3 * j[1];
    REAL kz     = (REAL)3 * j[2];
    REAL kxSqr_ = invRho / (REAL)2 * kx*kx;
    REAL kySqr_ = invRho / (REAL)2 * ky*ky;
    REAL kzSqr_ = invRho / (REAL)2 * kz*kz;
    REAL kxky_  = invRho * kx*ky;
    REAL kxkz_  = invRho * kx*kz;
    REAL kykz_  = invRho * ky*kz;
    REAL C1 = rhoBar + invRho*(REAL)3*jSqr;
    REAL C2, C3;
    C3 = -kxSqr_ - kySqr_ - kzSqr_;
    f[0] *= one_m_omega; f[0] += t0_omega * (C1+C3);
#define PLBO_PAIR(I, TW) f[I] *= one_m_omega; f[I] += TW * (C1+C2+C3); f[(I)+h] *= one_m_omega; f[(I)+h] += TW * (C1-C2+C3);
    C2 = -kx;       C3 = -kySqr_ - kzSqr_;   PLBO_PAIR(1, t1_omega)
    C2 = -ky;       C3 = -kxSqr_ - kzSqr_;   PLBO_PAIR(2, t1_omega)
    C2 = -kz;       C3 = -kxSqr_ - kySqr_;   PLBO_PAIR(3, t1_omega)
    C2 = -kx - ky;  C3 = kxky_ - kzSqr_;     PLBO_PAIR(4, t4_omega)
    C2 = -kx + ky;  C3 = -kxky_ - kzSqr_;    PLBO_PAIR(5, t4_omega)
    C2 = -kx - kz;  C3 = kxkz_ - kySqr_;     PLBO_PAIR(6, t4_omega)
    C2 = -kx + kz;  C3 = -kxkz_ - kySqr_;    PLBO_PAIR(7, t4_omega)
    C2 = -ky - kz;  C3 = kykz_ - kxSqr_;     PLBO_PAIR(8, t4_omega)
    C2 = -ky + kz;  C3 = -kykz_ - kxSqr_;    PLBO_PAIR(9, t4_omega)
    if (L->q == 27) {
        REAL t10_omega = L->t[10] * omega;
        C2 = -kx -ky -kz;  C3 = kxky_ + kxkz_ + kykz_;    PLBO_PAIR(10, t10_omega)
        C2 = -kx -ky +kz;  C3 = kxky_ - kxkz_ - kykz_;    PLBO_PAIR(11, t10_omega)
        C2 = -kx +ky -kz;  C3 = -kxky_ + kxkz_ - kykz_;   PLBO_PAIR(12, t10_omega)
        C2 = -kx +ky +kz;  C3 = -kxky_ - kxkz_ + kykz_;   PLBO_PAIR(13, t10_omega)
    }
#undef PLBO_PAIR
    return invRho*invRho*jSqr;
}

/* neqPiD3Q19 dynamicsTemplates3D.h:1903-1975 and neqPiD3Q27 :38-135: fromPiToFneq0..9 / 0..13.
 * PiNeq index: xx=0 xy=1 xz=2 yy=3 yz=4 zz=5. */
static void SFX(neqPi)(int q, const REAL* pi, REAL* n) {
    const REAL xx = pi[0], xy = pi[1], xz = pi[2], yy = pi[3], yz = pi[4], zz = pi[5];
    if (q == 19) {
        n[0] = (REAL)3./(REAL)2 * ( -(REAL)1./(REAL)3*xx - (REAL)1./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[1] = (REAL)1./(REAL)4 * (  (REAL)2./(REAL)3*xx - (REAL)1./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[2] = (REAL)1./(REAL)4 * ( -(REAL)1./(REAL)3*xx + (REAL)2./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[3] = (REAL)1./(REAL)4 * ( -(REAL)1./(REAL)3*xx - (REAL)1./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[4] = (REAL)1./(REAL)8 * (  (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy - (REAL)1./(REAL)3*zz + (REAL)2*xy );
        n[5] = (REAL)1./(REAL)8 * (  (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy - (REAL)1./(REAL)3*zz - (REAL)2*xy );
        n[6] = (REAL)1./(REAL)8 * (  (REAL)2./(REAL)3*xx - (REAL)1./(REAL)3*yy + (REAL)2./(REAL)3*zz + (REAL)2*xz );
        n[7] = (REAL)1./(REAL)8 * (  (REAL)2./(REAL)3*xx - (REAL)1./(REAL)3*yy + (REAL)2./(REAL)3*zz - (REAL)2*xz );
        n[8] = (REAL)1./(REAL)8 * ( -(REAL)1./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz + (REAL)2*yz );
        n[9] = (REAL)1./(REAL)8 * ( -(REAL)1./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz - (REAL)2*yz );
    } else {
        n[0] = (REAL)4./(REAL)3 * ( -(REAL)1./(REAL)3*xx - (REAL)1./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[1] = (REAL)1./(REAL)3 * (  (REAL)2./(REAL)3*xx - (REAL)1./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[2] = (REAL)1./(REAL)3 * ( -(REAL)1./(REAL)3*xx + (REAL)2./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[3] = (REAL)1./(REAL)3 * ( -(REAL)1./(REAL)3*xx - (REAL)1./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[4] = (REAL)1./(REAL)12 * (   (REAL)2 * xy + (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[5] = (REAL)1./(REAL)12 * ( - (REAL)2 * xy + (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy - (REAL)1./(REAL)3*zz );
        n[6] = (REAL)1./(REAL)12 * (   (REAL)2 * xz + (REAL)2./(REAL)3*xx - (REAL)1./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[7] = (REAL)1./(REAL)12 * ( - (REAL)2 * xz + (REAL)2./(REAL)3*xx - (REAL)1./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[8] = (REAL)1./(REAL)12 * (   (REAL)2 * yz - (REAL)1./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[9] = (REAL)1./(REAL)12 * ( - (REAL)2 * yz - (REAL)1./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[10] = (REAL)1./(REAL)48 * ( + (REAL)2*xy + (REAL)2*xz + (REAL)2*yz
                                      + (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[11] = (REAL)1./(REAL)48 * ( + (REAL)2*xy - (REAL)2*xz - (REAL)2*yz
                                      + (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[12] = (REAL)1./(REAL)48 * ( - (REAL)2*xy + (REAL)2*xz - (REAL)2*yz
                                      + (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz );
        n[13] = (REAL)1./(REAL)48 * ( - (REAL)2*xy - (REAL)2*xz + (REAL)2*yz
                                      + (REAL)2./(REAL)3*xx + (REAL)2./(REAL)3*yy + (REAL)2./(REAL)3*zz );
    }
}

/* dynamicsTemplates3D.h:2500-2554 (D3Q19 rlb_collision), :1683-1760 (D3Q27) */
static REAL SFX(rlb_collision)(const SFX(Lat)* L, REAL* f, REAL rhoBar, REAL invRho, const REAL j[3], const REAL PiNeq[6], REAL omega) {
    const int h = L->q / 2;
    const REAL jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
    REAL n[14];
    int i;
    SFX(neqPi)(L->q, PiNeq, n);
    f[0] = SFX(bgk_ma2_equilibrium)(L, 0, rhoBar, invRho, j, jSqr) + ((REAL)1-omega) * n[0];
    for (i = 1; i <= h; ++i)
        f[i] = SFX(bgk_ma2_equilibrium)(L, i, rhoBar, invRho, j, jSqr) + ((REAL)1-omega) * n[i];
    for (i = 1; i <= h; ++i)
        f[i+h] = SFX(bgk_ma2_equilibrium)(L, i+h, rhoBar, invRho, j, jSqr) + ((REAL)1-omega) * n[i];
    return jSqr*invRho*invRho;
}

/* complexDynamics/smagorinskyDynamics.hh:45-59 (SmagoOperations) + tensorNormSqr
 * geometricOperationTemplates.h:254-257 */
static REAL SFX(smago_prefactor)(REAL omega0, REAL cSmago, REAL invCs2) {
    REAL x = cSmago*omega0*invCs2;
    return (REAL)0.5 * (x*x);                       /* util::sqr(x) = x*x */
}
static REAL SFX(smago_omega)(REAL omega0, REAL preFactor, REAL rhoBar, const REAL m[6]) {
    REAL PiNeqNormSqr = m[0]*m[0]+m[3]*m[3]+m[5]*m[5]
                        + (REAL)2*m[1]*m[1] + (REAL)2*m[2]*m[2] + (REAL)2*m[4]*m[4];
    REAL PiNeqNorm    = SQRT(PiNeqNormSqr);
    REAL alpha        = preFactor * INVRHO(rhoBar);
    REAL linearTerm   = alpha*PiNeqNorm;
    REAL squareTerm   = (REAL)2*alpha*alpha*PiNeqNormSqr;
    return omega0*(1-linearTerm+squareTerm);
}

/* Cell::collide -> Dynamics::collide (core/cell.h:183-186).  Returns 1 and fills
 * rhoBar/uSqr when the dynamics gathers statistics. */
static int SFX(collide_cell)(const SFX(Lat)* L, const SFX(Dyn)* d, REAL* f, REAL* rhoBarOut, REAL* uSqrOut) {
    const int q = L->q, h = q / 2;
    REAL rhoBar, j[3], PiNeq[6], uSqr;
    int i;
    switch (d->type) {
    case PLBO_DYN_BGK:            /* basicDynamics/isoThermalDynamics.hh:319-331 */
        SFX(get_rhoBar_j)(q, f, &rhoBar, j);
        uSqr = SFX(bgk_ma2_collision)(L, f, rhoBar, j, d->omega);
        break;
    case PLBO_DYN_RLB: {          /* isoThermalDynamics.hh:1588-1602 */
        REAL invRho;
        SFX(rhoBar_j_PiNeq)(q, f, L->cs2, &rhoBar, j, PiNeq);
        invRho = INVRHO(rhoBar);
        uSqr = SFX(rlb_collision)(L, f, rhoBar, invRho, j, PiNeq, d->omega);
        break; }
    case PLBO_DYN_TRT: {          /* complexDynamics/trtDynamics.hh:64-99, sMinus = 1.1 (:40) */
        const REAL sPlus = d->omega;
        const REAL sMinus = (REAL)1.1;
        REAL eq[27], eq_plus[14], eq_minus[14], f_plus[14], f_minus[14], jSqr, invRho;
        SFX(get_rhoBar_j)(q, f, &rhoBar, j);
        jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
        invRho = INVRHO(rhoBar);
        SFX(bgk_ma2_equilibria)(L, rhoBar, invRho, j, jSqr, eq);
        for (i = 1; i <= h; ++i) {
            eq_plus[i]  = 0.5*(eq[i] + eq[i+h]);
            eq_minus[i] = 0.5*(eq[i] - eq[i+h]);
            f_plus[i]   = 0.5*(f[i] + f[i+h]);
            f_minus[i]  = 0.5*(f[i] - f[i+h]);
        }
        f[0] += -sPlus*f[0] + sPlus*eq[0];
        for (i = 1; i <= h; ++i) {
            f[i]   += -sPlus*(f_plus[i]-eq_plus[i]) - sMinus*(f_minus[i]-eq_minus[i]);
            f[i+h] += -sPlus*(f_plus[i]-eq_plus[i]) + sMinus*(f_minus[i]-eq_minus[i]);
        }
        uSqr = jSqr * invRho * invRho;
        break; }
    case PLBO_DYN_SMAGO: {        /* smagorinskyDynamics.hh:290-304 */
        REAL omega;
        SFX(rhoBar_j_PiNeq)(q, f, L->cs2, &rhoBar, j, PiNeq);
        omega = SFX(smago_omega)(d->omega, d->preFactor, rhoBar, PiNeq);
        uSqr = SFX(bgk_ma2_collision)(L, f, rhoBar, j, omega);
        break; }
    case PLBO_DYN_INCBGK:         /* IncBGKdynamics::collide isoThermalDynamics.hh:1405-1417, invRho0 = 1 */
        SFX(get_rhoBar_j)(q, f, &rhoBar, j);
        uSqr = SFX(bgk_ma2_collision_base)(L, f, rhoBar, j, d->omega, (REAL)1);
        break;
    case PLBO_DYN_BOUNCEBACK:     /* core/dynamics.hh:783-791 */
        for (i = 1; i <= h; ++i) { REAL tmp = f[i]; f[i] = f[i+h]; f[i+h] = tmp; }
        return 0;
    default:                      /* NoDynamics, core/dynamics.hh:968-972 */
        return 0;
    }
    *rhoBarOut = rhoBar; *uSqrOut = uSqr;
    return 1;
}

/* Dynamics::collideExternal(cell, rhoBar, j, thetaBar, stat) with moments taken from fields (basicDynamics/
 * dynamicsProcessor3D.hh:41-76): BGK isoThermalDynamics.hh:333-338, IncBGK :1422-1427, RegularizedBGK :1605-1615,
 * TRT complexDynamics/trtDynamics.hh:102-134, SmagorinskyBGK smagorinskyDynamics.hh:306-320, BounceBack
 * core/dynamics.hh:793-799, NoDynamics :974-978.  (Statistics are not restated: the fork's application never evaluates
 * them in this mode -- executeInternalProcessors() + incrementTime(), viscosityTest/src/test.cpp:64-65.) */
static void SFX(collide_external_cell)(const SFX(Lat)* L, const SFX(Dyn)* d, REAL* f, REAL rhoBar, const REAL j[3]) {
    const int q = L->q, h = q / 2;
    REAL PiNeq[6];
    int i;
    switch (d->type) {
    case PLBO_DYN_BGK:
        SFX(bgk_ma2_collision)(L, f, rhoBar, j, d->omega);
        break;
    case PLBO_DYN_INCBGK:
        SFX(bgk_ma2_collision_base)(L, f, rhoBar, j, d->omega, (REAL)1);
        break;
    case PLBO_DYN_RLB: {
        REAL invRho;
        SFX(compute_PiNeq)(q, f, L->cs2, rhoBar, j, PiNeq);
        invRho = INVRHO(rhoBar);
        SFX(rlb_collision)(L, f, rhoBar, invRho, j, PiNeq, d->omega);
        break; }
    case PLBO_DYN_TRT: {
        const REAL sPlus = d->omega;
        const REAL sMinus = (REAL)1.1;
        REAL eq[27], eq_plus[14], eq_minus[14], f_plus[14], f_minus[14], jSqr, invRho;
        jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
        invRho = INVRHO(rhoBar);
        SFX(bgk_ma2_equilibria)(L, rhoBar, invRho, j, jSqr, eq);
        for (i = 1; i <= h; ++i) {
            eq_plus[i]  = 0.5*(eq[i] + eq[i+h]);
            eq_minus[i] = 0.5*(eq[i] - eq[i+h]);
            f_plus[i]   = 0.5*(f[i] + f[i+h]);
            f_minus[i]  = 0.5*(f[i] - f[i+h]);
        }
        f[0] += -sPlus*f[0] + sPlus*eq[0];
        for (i = 1; i <= h; ++i) {
            f[i]   += -sPlus*(f_plus[i]-eq_plus[i]) - sMinus*(f_minus[i]-eq_minus[i]);
            f[i+h] += -sPlus*(f_plus[i]-eq_plus[i]) + sMinus*(f_minus[i]-eq_minus[i]);
        }
        break; }
    case PLBO_DYN_SMAGO: {
        REAL omega;
        SFX(compute_PiNeq)(q, f, L->cs2, rhoBar, j, PiNeq);
        omega = SFX(smago_omega)(d->omega, d->preFactor, rhoBar, PiNeq);
        SFX(bgk_ma2_collision)(L, f, rhoBar, j, omega);
        break; }
    case PLBO_DYN_BOUNCEBACK:
        for (i = 1; i <= h; ++i) { REAL tmp = f[i]; f[i] = f[i+h]; f[i+h] = tmp; }
        break;
    default:
        break;
    }
}

/* ------------------------------------------------------------- regularization ---- */

/* latticeBoltzmann/offEquilibriumTemplates.h:91-114 (generic; no 3D specialisation exists) */
static REAL SFX(fromPiToFneq)(const SFX(Lat)* L, int i, const REAL PiNeq[6]) {
    REAL fNeq = (REAL)0;
    int iPi = 0, a, b;
    for (a = 0; a < 3; ++a) {
        fNeq += PiNeq[iPi] * (L->c[i][a]*L->c[i][a] - L->cs2);
        ++iPi;
        for (b = a+1; b < 3; ++b) {
            fNeq += PiNeq[iPi] * (REAL)2 * L->c[i][a]*L->c[i][b];
            ++iPi;
        }
    }
    fNeq *= L->t[i] * L->invCs2 * L->invCs2 / (REAL)2;
    return fNeq;
}

/* IsoThermalBulkDynamics::regularize basicDynamics/isoThermalDynamics.hh:54-70;
 * BounceBack/NoDynamics::regularize are no-ops (core/dynamics.hh:807-811, :986-990) */
static void SFX(regularize)(const SFX(Lat)* L, const SFX(Dyn)* d, REAL* f, REAL rhoBar, const REAL j[3], REAL jSqr, const REAL PiNeq[6]) {
    const int h = L->q / 2;
    int i;
    if (d->type == PLBO_DYN_BOUNCEBACK || d->type == PLBO_DYN_NODYNAMICS) return;
    f[0] = SFX(computeEquilibrium)(L, d, 0, rhoBar, j, jSqr) + SFX(fromPiToFneq)(L, 0, PiNeq);
    for (i = 1; i <= h; ++i) {
        REAL fNeq;
        f[i]   = SFX(computeEquilibrium)(L, d, i, rhoBar, j, jSqr);
        f[i+h] = SFX(computeEquilibrium)(L, d, i+h, rhoBar, j, jSqr);
        fNeq = SFX(fromPiToFneq)(L, i, PiNeq);
        f[i]   += fNeq;
        f[i+h] += fNeq;
    }
}

/* ------------------------------------------------- plane boundary dynamics ---- */

/* sums over indexTemplates::subIndex<D,direction,value>() in ascending index order
 * (latticeBoltzmann/indexTemplates.h:78-99) */
static REAL SFX(sub_sum)(const SFX(Lat)* L, const REAL* f, int dir, int value) {
    REAL s = (REAL)0;
    int i;
    for (i = 0; i < L->q; ++i) if (L->c[i][dir] == value) s += f[i];
    return s;
}

/* VelocityDirichletBoundaryDynamics::computeRhoBar boundaryCondition/boundaryDynamics.hh:639-668
 * (no external force in these descriptors: the 0.5*force terms vanish) */
/* velIsJ: Dynamics::velIsJ of the base dynamics (true for IncBGKdynamics isoThermalDynamics.hh:1437-1440; composite
 * dynamics forward it, core/dynamics.hh:424-426) */
#define VEL_IS_J(d) ((d)->type == PLBO_DYN_INCBGK)
static REAL SFX(veldir_rhoBar)(const SFX(Lat)* L, const SFX(Bc)* b, const REAL* f, int velIsJ) {
    REAL rhoOnWall = SFX(sub_sum)(L, f, b->dir, 0);
    REAL rhoNormal = SFX(sub_sum)(L, f, b->dir, b->ori);
    REAL velNormal = (REAL)b->ori * ( b->u[b->dir]-0.5*(REAL)0 );
    if (velIsJ) return (REAL)2*rhoNormal+rhoOnWall-velNormal;
    return ((REAL)2*rhoNormal+rhoOnWall-1*velNormal) / ((REAL)1+velNormal);
}

/* DensityDirichletBoundaryDynamics::computeJ boundaryDynamics.hh:760-787 */
static void SFX(rhodir_j)(const SFX(Lat)* L, const SFX(Bc)* b, const REAL* f, REAL j[3]) {
    REAL rhoOnWall = SFX(sub_sum)(L, f, b->dir, 0);
    REAL rhoNormal = SFX(sub_sum)(L, f, b->dir, b->ori);
    j[0] = j[1] = j[2] = (REAL)0;
    j[b->dir] = (REAL)b->ori*((REAL)2*rhoNormal+rhoOnWall-b->rhoBar);
}

/* computeRhoBarJ of the plane boundary dynamics: velocity boundaryDynamics.hh:680-701,
 * density :822-828 */
static void SFX(plane_rhoBar_j)(const SFX(Lat)* L, const SFX(Bc)* b, const REAL* f, REAL* rhoBar, REAL j[3], int velIsJ) {
    if (b->kind == BC_VEL_REG || b->kind == BC_VEL_ZH) {
        REAL rho; int d;
        *rhoBar = SFX(veldir_rhoBar)(L, b, f, velIsJ);
        rho = FULLRHO(*rhoBar);
        for (d = 0; d < 3; ++d) j[d] = velIsJ ? (b->u[d] - 0.5*(REAL)0) : rho * (b->u[d] - 0.5*(REAL)0);
    } else {
        *rhoBar = b->rhoBar;
        SFX(rhodir_j)(L, b, f, j);
    }
}

/* RegularizedVelocityInnerEdgeDynamics3D<plane,normal1,normal2> / RegularizedVelocityInnerCornerDynamics3D<xN,yN,zN>
 * (boundaryCondition/regularizedBoundaryDynamics3D.h:40-140, .hh:40-298): composites of two / three velocity-Dirichlet plane
 * dynamics that share the stored velocity.  Bc encoding: edge  dir = plane, ori = 2*(normal1 > 0) + (normal2 > 0), with
 * direction1 = (plane+1)%3, direction2 = (plane+2)%3;  corner  ori = 4*(xN > 0) + 2*(yN > 0) + (zN > 0). */
static int SFX(inner_parts)(const SFX(Bc)* b, SFX(Bc) part[3]) {
    int n, k;
    if (b->kind == BC_INNER_EDGE) {
        n = 2;
        part[0] = *b; part[0].kind = BC_VEL_REG; part[0].dir = (b->dir + 1) % 3; part[0].ori = (b->ori & 2) ? 1 : -1;
        part[1] = *b; part[1].kind = BC_VEL_REG; part[1].dir = (b->dir + 2) % 3; part[1].ori = (b->ori & 1) ? 1 : -1;
    } else {
        n = 3;
        for (k = 0; k < 3; ++k) { part[k] = *b; part[k].kind = BC_VEL_REG; part[k].dir = k; part[k].ori = (b->ori & (4 >> k)) ? 1 : -1; }
    }
    return n;
}
/* computeRhoBar: (rhoBar1 + rhoBar2) / 2, (x + y + z) / 3   regularizedBoundaryDynamics3D.hh:95-101, :237-245 */
static REAL SFX(inner_rhoBar)(const SFX(Lat)* L, const SFX(Bc)* b, const REAL* f, int velIsJ) {
    SFX(Bc) part[3];
    int n = SFX(inner_parts)(b, part);
    if (n == 2) return (SFX(veldir_rhoBar)(L, &part[0], f, velIsJ) + SFX(veldir_rhoBar)(L, &part[1], f, velIsJ) ) / (REAL)2;
    return (SFX(veldir_rhoBar)(L, &part[0], f, velIsJ) +
            SFX(veldir_rhoBar)(L, &part[1], f, velIsJ) +
            SFX(veldir_rhoBar)(L, &part[2], f, velIsJ)) / (REAL)3;
}
/* computeDensity: the mean of the parts' densities  :86-92, :225-234 */
static REAL SFX(inner_density)(const SFX(Lat)* L, const SFX(Bc)* b, const REAL* f, int velIsJ) {
    SFX(Bc) part[3];
    int n = SFX(inner_parts)(b, part);
    if (n == 2) return (FULLRHO(SFX(veldir_rhoBar)(L, &part[0], f, velIsJ)) + FULLRHO(SFX(veldir_rhoBar)(L, &part[1], f, velIsJ)) ) / (REAL)2;
    return (FULLRHO(SFX(veldir_rhoBar)(L, &part[0], f, velIsJ)) +
            FULLRHO(SFX(veldir_rhoBar)(L, &part[1], f, velIsJ)) +
            FULLRHO(SFX(veldir_rhoBar)(L, &part[2], f, velIsJ))) / (REAL)3;
}
/* computeRhoBarJ: rhoBar as above, j = velocity (of the first part: all parts store the same), times rho unless velIsJ  :104-116, :248-260 */
static void SFX(inner_rhoBar_j)(const SFX(Lat)* L, const SFX(Bc)* b, const REAL* f, REAL* rhoBar, REAL j[3], int velIsJ) {
    REAL rho; int d;
    *rhoBar = SFX(inner_rhoBar)(L, b, f, velIsJ);
    rho = FULLRHO(*rhoBar);
    for (d = 0; d < 3; ++d) { j[d] = b->u[d]; if (!velIsJ) j[d] *= rho; }
}

/* boundaryTemplates<..>::compute_PiNeq boundaryCondition/boundaryTemplates.h:45-87 */
static void SFX(boundary_PiNeq)(const SFX(Lat)* L, const SFX(Dyn)* base, const SFX(Bc)* b, const REAL* f,
                                REAL rhoBar, const REAL j[3], REAL jSqr, REAL PiNeq[6])
{
    REAL fNeq[27];
    int i, iPi = 0, a, bb;
    for (i = 0; i < L->q; ++i) if (L->c[i][b->dir] == 0)
        fNeq[i] = f[i] - SFX(computeEquilibrium)(L, base, i, rhoBar, j, jSqr);
    for (i = 0; i < L->q; ++i) if (L->c[i][b->dir] == b->ori)
        fNeq[i] = (i == 0) ? (REAL)0 : f[i] - SFX(computeEquilibrium)(L, base, i, rhoBar, j, jSqr);
    for (a = 0; a < 3; ++a) {
        for (bb = a; bb < 3; ++bb) {
            PiNeq[iPi] = (REAL)0;
            for (i = 0; i < L->q; ++i) if (L->c[i][b->dir] == 0)
                PiNeq[iPi] += L->c[i][a]*L->c[i][bb]*fNeq[i];
            for (i = 0; i < L->q; ++i) if (L->c[i][b->dir] == b->ori)
                PiNeq[iPi] += (REAL)2 * L->c[i][a]*L->c[i][bb]* fNeq[i];
            ++iPi;
        }
    }
}

/* ZouHeClosure boundaryCondition/zouHeDynamics.hh:43-115 */
static void SFX(zouhe_closure)(const SFX(Lat)* L, const SFX(Dyn)* base, const SFX(Bc)* b, REAL* f) {
    const int q = L->q, h = q / 2;
    int missing[9], diag[9], nMissing = 0, nDiag = 0, i, k, d, removed = 0;
    REAL rhoBar, j[3], jSqr, falseRhoBar, falseJ[3], falseRho, falseU[3], invRho, diff[3];
    /* subIndexOutgoing<L,direction,orientation>: opposite of subIndex<L,direction,orientation> */
    for (i = 0; i < q; ++i) if (L->c[i][b->dir] == b->ori)
        missing[nMissing++] = (i == 0) ? 0 : (i <= h ? i + h : i - h);
    for (k = 0; k < nMissing; ++k) {
        int nn = abs(L->c[missing[k]][0]) + abs(L->c[missing[k]][1]) + abs(L->c[missing[k]][2]);
        if (nn == 1 && !removed) { removed = 1; continue; }   /* erase the first axis-aligned one */
        diag[nDiag++] = missing[k];
    }
    SFX(plane_rhoBar_j)(L, b, f, &rhoBar, j, VEL_IS_J(base));
    jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
    for (k = 0; k < nMissing; ++k) {
        int pop = missing[k];
        int opp = (pop == 0) ? 0 : (pop <= h ? pop + h : pop - h);
        f[pop] = f[opp]
            - SFX(computeEquilibrium)(L, base, opp, rhoBar, j, jSqr)
            + SFX(computeEquilibrium)(L, base, pop, rhoBar, j, jSqr);
    }
    /* momentTemplates::compute_rho_uLb momentTemplates3D.h:437-445 */
    SFX(get_rhoBar_j)(q, f, &falseRhoBar, falseJ);
    invRho = INVRHO(falseRhoBar);
    falseRho = FULLRHO(falseRhoBar);
    falseU[0] = falseJ[0]; falseU[1] = falseJ[1]; falseU[2] = falseJ[2];
    falseU[0] *= invRho; falseU[1] *= invRho; falseU[2] *= invRho;
    for (d = 0; d < 3; ++d) diff[d] = (j[d] - falseRho*falseU[d]) / (REAL)nDiag;
    for (k = 0; k < nDiag; ++k)
        for (d = 1; d < 3; ++d)
            f[diag[k]] += L->c[diag[k]][(b->dir+d)%3] * diff[(b->dir+d)%3];
}

/* completePopulations of the four plane boundary dynamics:
 * RegularizedVelocity/DensityBoundaryDynamics regularizedBoundaryDynamics.hh:94-109, :166-181;
 * ZouHeVelocity/PressureDynamics zouHeDynamics.hh:166-169, :222-225 */
static void SFX(complete_populations)(const SFX(Lat)* L, const SFX(Dyn)* base, const SFX(Bc)* b, REAL* f) {
    if (b->kind == BC_VEL_REG || b->kind == BC_RHO_REG) {
        REAL rhoBar, j[3], jSqr, PiNeq[6];
        SFX(plane_rhoBar_j)(L, b, f, &rhoBar, j, VEL_IS_J(base));
        jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
        SFX(boundary_PiNeq)(L, base, b, f, rhoBar, j, jSqr, PiNeq);
        SFX(regularize)(L, base, f, rhoBar, j, jSqr, PiNeq);
    } else if (b->kind == BC_VEL_ZH || b->kind == BC_RHO_ZH) {
        SFX(zouhe_closure)(L, base, b, f);
    } else if (b->kind == BC_INNER_EDGE || b->kind == BC_INNER_CORNER) {
        /* regularizedBoundaryDynamics3D.hh:138-162 (edge), :276-298 (corner): bounce-back of off-equilibrium on the unknown
         * populations -- those that leave the solid along both (all three) normals --, then regularize ALL populations with
         * the base dynamics' PiNeq of the cell (IsoThermalBulkDynamics::computePiNeq isoThermalDynamics.hh:78-85: moments of
         * the populations themselves) around the boundary's own rhoBar and j */
        SFX(Bc) part[3];
        REAL rhoBar, j[3], jSqr, cellRhoBar, cellJ[3], PiNeq[6];
        int n = SFX(inner_parts)(b, part), i, k;
        SFX(inner_rhoBar_j)(L, b, f, &rhoBar, j, VEL_IS_J(base));
        jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
        for (i = 0; i < L->q; ++i) {
            int unknown = 1;
            for (k = 0; k < n; ++k) if (L->c[i][part[k].dir] != -part[k].ori) unknown = 0;
            if (unknown) {
                int h = (L->q - 1) / 2;
                int opp = (i == 0) ? 0 : (i <= h ? i + h : i - h);
                f[i] = f[opp]
                    - SFX(computeEquilibrium)(L, base, opp, rhoBar, j, jSqr)
                    + SFX(computeEquilibrium)(L, base, i, rhoBar, j, jSqr);
            }
        }
        if (base->type == PLBO_DYN_BOUNCEBACK || base->type == PLBO_DYN_NODYNAMICS) {
            for (i = 0; i < 6; ++i) PiNeq[i] = (REAL)0;            /* core/dynamics.hh:837-842, :1017-1022 */
        } else {
            SFX(get_rhoBar_j)(L->q, f, &cellRhoBar, cellJ);
            SFX(compute_PiNeq)(L->q, f, L->cs2, cellRhoBar, cellJ, PiNeq);
        }
        SFX(regularize)(L, base, f, rhoBar, j, jSqr, PiNeq);
    }
    /* BC_STORE_VEL: StoreVelocityDynamics inherits the empty completePopulations
     * (boundaryDynamics.hh:185-187) */
}

/* ------------------------------------------- Cell::computeDensity/Velocity/RhoBar ---- */

static const SFX(Dyn)* SFX(dyn_of)(const SFX(Lat)* L, size_t cell) { return &L->dyns[L->dynIdx[cell]]; }

/* Dynamics::computeDensity of whatever sits on the cell:
 *  bulk  BasicBulkDynamics::computeDensity core/dynamics.hh:277-280 (compute_rho = fullRho(get_rhoBar))
 *  BB/NoDynamics return their rho                       core/dynamics.hh:813-815, :992-994
 *  velocity planes VelocityDirichlet..::computeDensity  boundaryDynamics.hh:671-677
 *  density planes  StoreDensityDynamics::computeDensity boundaryDynamics.hh:247-250
 *  edges/corners   BoundaryCompositeDynamics::computeDensity -> base (:83-87) */
static REAL SFX(cell_density)(const SFX(Lat)* L, size_t cell) {
    const REAL* f = L->f + cell * L->q;
    const SFX(Dyn)* d = SFX(dyn_of)(L, cell);
    int bi = L->bcIdx[cell];
    if (bi >= 0) {
        const SFX(Bc)* b = &L->bcs[bi];
        if (b->kind == BC_VEL_REG || b->kind == BC_VEL_ZH) return FULLRHO(SFX(veldir_rhoBar)(L, b, f, VEL_IS_J(d)));
        if (b->kind == BC_RHO_REG || b->kind == BC_RHO_ZH) return FULLRHO(b->rhoBar);
        if (b->kind == BC_INNER_EDGE || b->kind == BC_INNER_CORNER) return SFX(inner_density)(L, b, f, VEL_IS_J(d));
    }
    if (d->type == PLBO_DYN_BOUNCEBACK || d->type == PLBO_DYN_NODYNAMICS) return d->rho;
    return FULLRHO(SFX(get_rhoBar)(L->q, f));
}

/* Dynamics::computeVelocity:
 *  bulk  compute_uLb momentTemplates3D.h:429-435;  BB/NoDynamics zero (core/dynamics.hh:821-826)
 *  StoreVelocityDynamics (velocity planes, edges, corners) boundaryDynamics.hh:324-331
 *  density planes DensityDirichlet..::computeVelocity boundaryDynamics.hh:790-812 */
static void SFX(cell_velocity)(const SFX(Lat)* L, size_t cell, REAL u[3]) {
    const REAL* f = L->f + cell * L->q;
    const SFX(Dyn)* d = SFX(dyn_of)(L, cell);
    int bi = L->bcIdx[cell];
    if (bi >= 0) {
        const SFX(Bc)* b = &L->bcs[bi];
        if (b->kind == BC_RHO_REG || b->kind == BC_RHO_ZH) {
            REAL invRho = INVRHO(b->rhoBar);
            int k;
            SFX(rhodir_j)(L, b, f, u);
            for (k = 0; k < 3; ++k) { if (!VEL_IS_J(d)) u[k] *= invRho; u[k] += 0.5*(REAL)0; }
            return;
        }
        u[0] = b->u[0]; u[1] = b->u[1]; u[2] = b->u[2];
        return;
    }
    if (d->type == PLBO_DYN_BOUNCEBACK || d->type == PLBO_DYN_NODYNAMICS) { u[0] = u[1] = u[2] = (REAL)0; return; }
    if (d->type == PLBO_DYN_INCBGK) {   /* IncBGKdynamics::computeVelocity isoThermalDynamics.hh:1389-1396: u = j * invRho0 */
        REAL dummyRhoBar;
        SFX(get_rhoBar_j)(L->q, f, &dummyRhoBar, u);
        u[0] *= (REAL)1; u[1] *= (REAL)1; u[2] *= (REAL)1;
        return;
    }
    {
        REAL invRho;
        SFX(get_j)(L->q, f, u);
        invRho = INVRHO(SFX(get_rhoBar)(L->q, f));
        u[0] *= invRho; u[1] *= invRho; u[2] *= invRho;
    }
}

/* Dynamics::computeRhoBar:
 *  bulk get_rhoBar core/dynamics.hh:351-354; BB rhoBar(rho) :861-863;
 *  velocity planes boundaryDynamics.hh:639-668; density planes :254-257;
 *  StoreVelocityDynamics (edges, corners) rhoBar(computeDensity) :342-346 */
static REAL SFX(cell_rhoBar)(const SFX(Lat)* L, size_t cell) {
    const REAL* f = L->f + cell * L->q;
    const SFX(Dyn)* d = SFX(dyn_of)(L, cell);
    int bi = L->bcIdx[cell];
    if (bi >= 0) {
        const SFX(Bc)* b = &L->bcs[bi];
        if (b->kind == BC_VEL_REG || b->kind == BC_VEL_ZH) return SFX(veldir_rhoBar)(L, b, f, VEL_IS_J(d));
        if (b->kind == BC_RHO_REG || b->kind == BC_RHO_ZH) return b->rhoBar;
        if (b->kind == BC_INNER_EDGE || b->kind == BC_INNER_CORNER) return SFX(inner_rhoBar)(L, b, f, VEL_IS_J(d));
        { REAL rho = SFX(cell_density)(L, cell); return RHOBAR(rho); }
    }
    if (d->type == PLBO_DYN_BOUNCEBACK || d->type == PLBO_DYN_NODYNAMICS) return RHOBAR(d->rho);
    return SFX(get_rhoBar)(L->q, f);
}

/* Dynamics::computeRhoBarJ of whatever sits on the cell (what BoxRhoBarJfunctional3D stores,
 * dataProcessors/dataAnalysisFunctional3D.hh:528-548):
 *  bulk get_rhoBar_j core/dynamics.hh:357-361; BounceBack rhoBar(rho), 0 :880-885; NoDynamics 1 - SkordosFactor, 0 :1061-1066;
 *  velocity planes boundaryDynamics.hh:682-701; density planes :816-822;
 *  StoreVelocityDynamics (edges, corners) rho = computeDensity, j = rho * velocity :348-359 */
static void SFX(cell_rhoBar_j)(const SFX(Lat)* L, size_t cell, REAL* rhoBar, REAL j[3]) {
    const REAL* f = L->f + cell * L->q;
    const SFX(Dyn)* d = SFX(dyn_of)(L, cell);
    int bi = L->bcIdx[cell], k;
    if (bi >= 0) {
        const SFX(Bc)* b = &L->bcs[bi];
        if (b->kind == BC_STORE_VEL) {
            REAL rho = SFX(cell_density)(L, cell);
            *rhoBar = RHOBAR(rho);
            for (k = 0; k < 3; ++k) j[k] = rho * (b->u[k] - 0.5*(REAL)0);
        } else if (b->kind == BC_INNER_EDGE || b->kind == BC_INNER_CORNER) {
            SFX(inner_rhoBar_j)(L, b, f, rhoBar, j, VEL_IS_J(d));
        } else {
            SFX(plane_rhoBar_j)(L, b, f, rhoBar, j, VEL_IS_J(d));
        }
        return;
    }
    if (d->type == PLBO_DYN_BOUNCEBACK) { *rhoBar = RHOBAR(d->rho); j[0] = j[1] = j[2] = (REAL)0; return; }
    if (d->type == PLBO_DYN_NODYNAMICS) { *rhoBar = (REAL)1 - (REAL)1; j[0] = j[1] = j[2] = (REAL)0; return; }
    SFX(get_rhoBar_j)(L->q, f, rhoBar, j);
}

/* BoxRhoBarJfunctional3D::processGenericBlocks dataAnalysisFunctional3D.hh:528-548 (cells of the bulk only) */
static void SFX(box_rhoBar_j)(SFX(Lat)* L, const int box[6]) {
    int x, y, z;
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z) {
        size_t cell = CELL(L, x, y, z);
        const int* b = L->statBox;
        if (x < b[0] || x > b[1] || y < b[2] || y > b[3] || z < b[4] || z > b[5]) continue;
        SFX(cell_rhoBar_j)(L, cell, &L->ext[4*cell], &L->ext[4*cell + 1]);
    }
}

/* ------------------------------------------------ non-local boundary processors ---- */

/* common tail of Outer{Edge,Corner}Functional3D::processCell
 * (finiteDifferenceBoundaryProcessor3D.hh:189-221, :287-313) */
static void SFX(fd_regularize)(SFX(Lat)* L, size_t cell, REAL rho, REAL dA_uB[3][3]) {
    const SFX(Dyn)* d = SFX(dyn_of)(L, cell);
    REAL omega = (d->type == PLBO_DYN_BOUNCEBACK || d->type == PLBO_DYN_NODYNAMICS) ? (REAL)0 : d->omega;
    REAL sToPi = - rho / L->invCs2 / omega;
    REAL pi[6], vel[3], j[3], jSqr;
    pi[0] = (REAL)2 * dA_uB[0][0] * sToPi;
    pi[3] = (REAL)2 * dA_uB[1][1] * sToPi;
    pi[5] = (REAL)2 * dA_uB[2][2] * sToPi;
    pi[1] = (dA_uB[0][1]+dA_uB[1][0]) * sToPi;
    pi[2] = (dA_uB[0][2]+dA_uB[2][0]) * sToPi;
    pi[4] = (dA_uB[1][2]+dA_uB[2][1]) * sToPi;
    SFX(cell_velocity)(L, cell, vel);
    if (VEL_IS_J(d)) { j[0] = vel[0]; j[1] = vel[1]; j[2] = vel[2]; }      /* finiteDifferenceBoundaryProcessor3D.hh:200-205 */
    else { j[0] = rho*vel[0]; j[1] = rho*vel[1]; j[2] = rho*vel[2]; }
    jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
    SFX(regularize)(L, d, L->f + cell * L->q, RHOBAR(rho), j, jSqr, pi);
}

/* fd::DirectedGradients3D<..,orthogonal=true>::o1_velocityDerivative finiteDifference3D.h:64-82 */
static void SFX(grad_normal)(const SFX(Lat)* L, int x, int y, int z, int dir, int ori, REAL out[3]) {
    REAL u0[3], u1[3];
    int k;
    SFX(cell_velocity)(L, CELL(L, x, y, z), u0);
    SFX(cell_velocity)(L, CELL(L, x + (dir==0 ? -ori : 0), y + (dir==1 ? -ori : 0), z + (dir==2 ? -ori : 0)), u1);
    for (k = 0; k < 3; ++k) out[k] = -ori * (-u0[k] + u1[k]);
}
/* ..orthogonal=false: o2_velocityDerivative, central difference finiteDifference3D.h:158-178 */
static void SFX(grad_tangent)(const SFX(Lat)* L, int x, int y, int z, int dir, REAL out[3]) {
    REAL up[3], um[3];
    int k;
    SFX(cell_velocity)(L, CELL(L, x + (dir==0), y + (dir==1), z + (dir==2)), up);
    SFX(cell_velocity)(L, CELL(L, x - (dir==0), y - (dir==1), z - (dir==2)), um);
    for (k = 0; k < 3; ++k) out[k] = (up[k] - um[k]) / (REAL)2;
}

/* OuterVelocityEdgeFunctional3D::processCell finiteDifferenceBoundaryProcessor3D.hh:164-226 */
static void SFX(edge_cell)(SFX(Lat)* L, int x, int y, int z, int plane, int n1, int n2) {
    const int d1 = (plane+1)%3, d2 = (plane+2)%3;
    int c10[3] = {x, y, z}, c01[3] = {x, y, z};
    REAL rho10, rho01, rho, dA_uB[3][3];
    c10[d1] += -n1;
    c01[d2] += -n2;
    rho10 = SFX(cell_density)(L, CELL(L, c10[0], c10[1], c10[2]));
    rho01 = SFX(cell_density)(L, CELL(L, c01[0], c01[1], c01[2]));
    rho = (REAL)0.5*(rho01+rho10);
    SFX(grad_tangent)(L, x, y, z, plane, dA_uB[plane]);
    SFX(grad_normal)(L, x, y, z, d1, n1, dA_uB[d1]);
    SFX(grad_normal)(L, x, y, z, d2, n2, dA_uB[d2]);
    SFX(fd_regularize)(L, CELL(L, x, y, z), rho, dA_uB);
}

/* PlaneFdBoundaryFunctional3D::processCell finiteDifferenceBoundaryProcessor3D.hh:68-117 (createInterpBoundaryCondition3D,
 * boundaryCondition3D.hh:1049-1085: Velocity/DensityDirichletBoundaryDynamics + this processor): density and velocity of the
 * cell itself through its Dirichlet dynamics, velocity gradients one-sided along the normal and central along the face,
 * then regularize */
static void SFX(plane_fd_cell)(SFX(Lat)* L, int x, int y, int z, int dir, int ori) {
    REAL rho, dA_uB[3][3];
    int d;
    rho = SFX(cell_density)(L, CELL(L, x, y, z));
    for (d = 0; d < 3; ++d) {
        if (d == dir) SFX(grad_normal)(L, x, y, z, d, ori, dA_uB[d]);
        else SFX(grad_tangent)(L, x, y, z, d, dA_uB[d]);
    }
    SFX(fd_regularize)(L, CELL(L, x, y, z), rho, dA_uB);
}

/* OuterVelocityCornerFunctional3D::processCell finiteDifferenceBoundaryProcessor3D.hh:264-318 */
static void SFX(corner_cell)(SFX(Lat)* L, int x, int y, int z, int xn, int yn, int zn) {
    REAL rho100 = SFX(cell_density)(L, CELL(L, x - xn, y, z));
    REAL rho010 = SFX(cell_density)(L, CELL(L, x, y - yn, z));
    REAL rho001 = SFX(cell_density)(L, CELL(L, x, y, z - zn));
    REAL rho = (REAL)1/(REAL)3 * (rho001 + rho010 + rho100);
    REAL dA_uB[3][3];
    SFX(grad_normal)(L, x, y, z, 0, xn, dA_uB[0]);
    SFX(grad_normal)(L, x, y, z, 1, yn, dA_uB[1]);
    SFX(grad_normal)(L, x, y, z, 2, zn, dA_uB[2]);
    SFX(fd_regularize)(L, CELL(L, x, y, z), rho, dA_uB);
}

static int SFX(in_stat_box)(const SFX(Lat)* L, int x, int y, int z) {
    const int* b = L->statBox;
    return x >= b[0] && x <= b[1] && y >= b[2] && y <= b[3] && z >= b[4] && z <= b[5];
}

/* AtomicBlock3D::executeInternalProcessors(level 0) atomicBlock/atomicBlock3D.cpp:195-222:
 * processors run in insertion order, each over its box in x,y,z order. */
static void SFX(execute_processors)(SFX(Lat)* L) {
    int p, x, y, z;
    for (p = 0; p < L->nProc; ++p) {
        const Proc* pr = &L->procs[p];
        for (x = pr->box[0]; x <= pr->box[1]; ++x)
        for (y = pr->box[2]; y <= pr->box[3]; ++y)
        for (z = pr->box[4]; z <= pr->box[5]; ++z) {
            size_t cell = CELL(L, x, y, z);
            if (!SFX(in_stat_box)(L, x, y, z)) continue;       /* envelope: processors act on the bulk only */
            if (pr->cls == PROC_PLANE) {
                /* WrappedLocalBoundaryFunctional3D::process wrappedLocalBoundaryProcessor3D.hh:42-62 */
                int bi = L->bcIdx[cell];
                if (bi >= 0) SFX(complete_populations)(L, SFX(dyn_of)(L, cell), &L->bcs[bi], L->f + cell * L->q);
            } else if (pr->cls == PROC_PLANE_FD) {
                /* PlaneFdBoundaryFunctional3D::processCell finiteDifferenceBoundaryProcessor3D.hh:68-117 has no isComposite() test:
                 * it runs on every cell of its box, also where defineDynamics has replaced the Dirichlet dynamics */
                SFX(plane_fd_cell)(L, x, y, z, pr->a, pr->b);
            } else if (pr->cls == PROC_EDGE) {
                SFX(edge_cell)(L, x, y, z, pr->a, pr->b, pr->c);
            } else {
                SFX(corner_cell)(L, x, y, z, pr->a, pr->b, pr->c);
            }
        }
    }
}

/* -------------------------------------------------------------------- the step ---- */

/* One MultiBlockLattice3D::collideAndStream() (multiBlock/multiBlockLattice3D.hh:318-360) for a
 * single block: BlockLattice3D::collideAndStream(domain) (atomicBlock/blockLattice3D.hh:253-312)
 * = collide every cell, then exact streaming in which a population that would leave the
 * sweep domain stays on its cell in the opposite slot (collide()+revert(), :173-185, followed by
 * the bounds-checked swaps of boundaryStream(), :403-427); a periodic direction behaves as a
 * wrap because the reference collides the periodic envelope copy and streams it in
 * (extendPeriodic :253-285 + duplicateOverlaps).  Then the level-0 processors, then
 * BlockStatistics::evaluate (core/blockStatistics.cpp:77-120). */
static void SFX(step)(SFX(Lat)* L) {
    const int q = L->q, h = q / 2, nx = L->nx, ny = L->ny, nz = L->nz;
    int x, y, z, i;
    REAL* tmp;
    for (x = 0; x < nx; ++x) for (y = 0; y < ny; ++y) for (z = 0; z < nz; ++z) {
        size_t cell = CELL(L, x, y, z);
        REAL rhoBar, uSqr;
        /* CompositeDynamics::collide core/dynamics.hh:470-475: a boundary dynamics that prepares the collision itself (the
         * dynamics-based managers) completes the populations first */
        if (L->bcIdx[cell] >= 0 && L->bcs[L->bcIdx[cell]].autoPrep)
            SFX(complete_populations)(L, SFX(dyn_of)(L, cell), &L->bcs[L->bcIdx[cell]], L->f + cell * q);
        if (SFX(collide_cell)(L, SFX(dyn_of)(L, cell), L->f + cell * q, &rhoBar, &uSqr) && SFX(in_stat_box)(L, x, y, z)) {
            /* gatherStatistics core/latticeStatistics.h:36-45 (always double) */
            L->tmpSumRhoBar += (double)rhoBar;
            L->tmpSumUSqr += (double)uSqr;
            if ((double)uSqr > L->tmpMaxUSqr) L->tmpMaxUSqr = (double)uSqr;
            L->tmpN++;
        }
    }
    for (x = 0; x < nx; ++x) for (y = 0; y < ny; ++y) for (z = 0; z < nz; ++z) {
        size_t cell = CELL(L, x, y, z);
        L->g[cell * q] = L->f[cell * q];
        for (i = 1; i < q; ++i) {
            int sx = x - L->c[i][0], sy = y - L->c[i][1], sz = z - L->c[i][2], ok = 1;
            if (sx < 0 || sx >= nx) { if (L->per[0]) sx = (sx + nx) % nx; else ok = 0; }
            if (sy < 0 || sy >= ny) { if (L->per[1]) sy = (sy + ny) % ny; else ok = 0; }
            if (sz < 0 || sz >= nz) { if (L->per[2]) sz = (sz + nz) % nz; else ok = 0; }
            if (ok) L->g[cell * q + i] = L->f[CELL(L, sx, sy, sz) * q + i];
            else    L->g[cell * q + i] = L->f[cell * q + (i <= h ? i + h : i - h)];
        }
    }
    tmp = L->f; L->f = L->g; L->g = tmp;
    SFX(execute_processors)(L);
    if (L->tmpN == 0) { L->avRhoBar = 0.; L->avUSqr = 0.; }
    else { L->avRhoBar = L->tmpSumRhoBar / (double)L->tmpN; L->avUSqr = L->tmpSumUSqr / (double)L->tmpN; }
    L->maxUSqr = L->tmpMaxUSqr; L->numCells = L->tmpN;
    L->tmpSumRhoBar = 0.; L->tmpSumUSqr = 0.; L->tmpMaxUSqr = -DBL_MAX; L->tmpN = 0;
}

/* One executeInternalProcessors() + incrementTime() of a lattice that carries ExternalRhoJcollideAndStream3D at level 0
 * and BoxRhoBarJfunctional3D at level 3 (src/variables.hh:485-486, viscosityTest/src/test.cpp:64-65).  The sweep
 * (basicDynamics/dynamicsProcessor3D.hh:108-195) works on the bulk ENLARGED by the vicinity: the whole array of this
 * oracle lattice, which the caller has created one cell larger on every non-periodic side (the reference's envelope;
 * statBox = the bulk).  Cells of that shell collide with the field values found there -- zero, nothing ever writes
 * them -- and stream like any other cell; on the rim of the enlarged box a population stays in the opposite slot
 * (collide()+revert() :41-57, bounds-checked boundaryStream :80-104).  Then the boundary condition's level-0 processors
 * (integrated after the sweep), then the moments of the post-stream populations into the fields. */
static void SFX(ext_step)(SFX(Lat)* L) {
    const int q = L->q, h = q / 2, nx = L->nx, ny = L->ny, nz = L->nz;
    int x, y, z, i;
    REAL* tmp;
    for (x = 0; x < nx; ++x) for (y = 0; y < ny; ++y) for (z = 0; z < nz; ++z) {
        size_t cell = CELL(L, x, y, z);
        if (L->bcIdx[cell] >= 0 && L->bcs[L->bcIdx[cell]].autoPrep)    /* CompositeDynamics::collideExternal core/dynamics.hh:478-484 */
            SFX(complete_populations)(L, SFX(dyn_of)(L, cell), &L->bcs[L->bcIdx[cell]], L->f + cell * q);
        SFX(collide_external_cell)(L, SFX(dyn_of)(L, cell), L->f + cell * q, L->ext[4*cell], &L->ext[4*cell + 1]);
    }
    for (x = 0; x < nx; ++x) for (y = 0; y < ny; ++y) for (z = 0; z < nz; ++z) {
        size_t cell = CELL(L, x, y, z);
        L->g[cell * q] = L->f[cell * q];
        for (i = 1; i < q; ++i) {
            int sx = x - L->c[i][0], sy = y - L->c[i][1], sz = z - L->c[i][2], ok = 1;
            if (sx < 0 || sx >= nx) { if (L->per[0]) sx = (sx + nx) % nx; else ok = 0; }
            if (sy < 0 || sy >= ny) { if (L->per[1]) sy = (sy + ny) % ny; else ok = 0; }
            if (sz < 0 || sz >= nz) { if (L->per[2]) sz = (sz + nz) % nz; else ok = 0; }
            if (ok) L->g[cell * q + i] = L->f[CELL(L, sx, sy, sz) * q + i];
            else    L->g[cell * q + i] = L->f[cell * q + (i <= h ? i + h : i - h)];
        }
    }
    tmp = L->f; L->f = L->g; L->g = tmp;
    SFX(execute_processors)(L);
    SFX(box_rhoBar_j)(L, L->statBox);
}

/* -------------------------------------------------------------------- set-up ---- */

static SFX(Lat)* SFX(create)(int q, int nx, int ny, int nz) {
    SFX(Lat)* L = (SFX(Lat)*)calloc(1, sizeof(SFX(Lat)));
    size_t n = (size_t)nx * ny * nz;
    L->q = q; L->nx = nx; L->ny = ny; L->nz = nz;
    SFX(init_descriptor)(L);
    L->f = (REAL*)calloc(n * q, sizeof(REAL));
    L->g = (REAL*)calloc(n * q, sizeof(REAL));
    L->dynIdx = (uint8_t*)calloc(n, 1);
    L->bcIdx = (int32_t*)malloc(n * sizeof(int32_t));
    { size_t k; for (k = 0; k < n; ++k) L->bcIdx[k] = -1; }
    L->statBox[0] = 0; L->statBox[1] = nx-1; L->statBox[2] = 0; L->statBox[3] = ny-1; L->statBox[4] = 0; L->statBox[5] = nz-1;
    L->tmpMaxUSqr = -DBL_MAX; L->maxUSqr = -DBL_MAX;
    return L;
}

static void SFX(destroy)(SFX(Lat)* L) {
    free(L->f); free(L->g); free(L->ext); free(L->dynIdx); free(L->bcIdx); free(L->bcs); free(L->procs); free(L);
}

static int SFX(find_dyn)(SFX(Lat)* L, int type, double p1, double p2) {
    SFX(Dyn) d;
    int k;
    memset(&d, 0, sizeof d);
    d.type = type;
    if (type == PLBO_DYN_BOUNCEBACK || type == PLBO_DYN_NODYNAMICS) d.rho = (REAL)p1;
    else d.omega = (REAL)p1;
    if (type == PLBO_DYN_SMAGO) { d.cSmago = (REAL)p2; d.preFactor = SFX(smago_prefactor)(d.omega, d.cSmago, L->invCs2); }
    for (k = 0; k < L->nDyn; ++k) if (memcmp(&L->dyns[k], &d, sizeof d) == 0) return k;
    if (L->nDyn >= MAX_DYN) return -1;
    L->dyns[L->nDyn] = d;
    return L->nDyn++;
}

static int SFX(add_bc)(SFX(Lat)* L, int kind, int dir, int ori, const int box[6]) {
    int x, y, z;
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z) {
        size_t cell = CELL(L, x, y, z);
        SFX(Bc)* b;
        if (L->bcIdx[cell] >= 0) return 0;            /* nested composites: not restated */
        if (L->nBc == L->capBc) {
            L->capBc = L->capBc ? 2 * L->capBc : 1024;
            L->bcs = (SFX(Bc)*)realloc(L->bcs, L->capBc * sizeof(SFX(Bc)));
        }
        b = &L->bcs[L->nBc];
        memset(b, 0, sizeof *b);
        b->kind = kind; b->dir = dir; b->ori = ori;
        b->rhoBar = RHOBAR((REAL)1);                  /* StoreDensityDynamics ctor boundaryDynamics.hh:200-205 */
        L->bcIdx[cell] = L->nBc++;
    }
    return 1;
}

static void SFX(mark_auto)(SFX(Lat)* L, const int box[6]) {
    int x, y, z;
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z) {
        size_t cell = CELL(L, x, y, z);
        if (L->bcIdx[cell] >= 0) L->bcs[L->bcIdx[cell]].autoPrep = 1;
    }
}

static void SFX(add_proc)(SFX(Lat)* L, int cls, const int box[6], int a, int b, int c) {
    Proc* p;
    if (L->nProc == L->capProc) {
        L->capProc = L->capProc ? 2 * L->capProc : 64;
        L->procs = (Proc*)realloc(L->procs, L->capProc * sizeof(Proc));
    }
    p = &L->procs[L->nProc++];
    p->cls = cls; memcpy(p->box, box, 6 * sizeof(int)); p->a = a; p->b = b; p->c = c;
}

/* iniCellAtEquilibrium core/cell.hh:124-132 */
static void SFX(ini_equilibrium)(SFX(Lat)* L, size_t cell, REAL density, const REAL velocity[3]) {
    REAL j[3], jSqr, rhoBar;
    int i;
    j[0] = velocity[0]*density; j[1] = velocity[1]*density; j[2] = velocity[2]*density;
    jSqr = j[0]*j[0] + j[1]*j[1] + j[2]*j[2];
    rhoBar = RHOBAR(density);
    for (i = 0; i < L->q; ++i)
        L->f[cell * L->q + i] = SFX(computeEquilibrium)(L, SFX(dyn_of)(L, cell), i, rhoBar, j, jSqr);
}

/* computeAverageEnergy dataAnalysisWrapper3D.hh:1252-1257 + BoxSumEnergyFunctional3D
 * dataAnalysisFunctional3D.hh:97-123; computeAverageDensity :1226-1230 + BoxSumRhoBar :62-76 */
static double SFX(average_energy)(const SFX(Lat)* L, const int box[6]) {
    double sum = 0.;
    int x, y, z;
    size_t n = (size_t)(box[1]-box[0]+1) * (box[3]-box[2]+1) * (box[5]-box[4]+1);
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z) {
        REAL u[3], uNormSqr;
        SFX(cell_velocity)(L, CELL(L, x, y, z), u);
        uNormSqr = u[0]*u[0] + u[1]*u[1] + u[2]*u[2];
        sum += (double)uNormSqr;
    }
    { REAL sumEnergy = (REAL)(sum / (REAL)2); return (double)(sumEnergy / (REAL)n); }
}
static double SFX(average_density)(const SFX(Lat)* L, const int box[6]) {
    double sum = 0.;
    int x, y, z;
    size_t n = (size_t)(box[1]-box[0]+1) * (box[3]-box[2]+1) * (box[5]-box[4]+1);
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z)
        sum += (double)SFX(cell_rhoBar)(L, CELL(L, x, y, z));
    { REAL s = (REAL)sum; return (double)FULLRHO(s / (REAL)n); }
}

#undef CELL
#undef RHOBAR
#undef FULLRHO
#undef INVRHO
