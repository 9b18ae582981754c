// palabos_b200/csrc/lb_core.cuh -- velocity sets and per-cell arithmetic of the hot path.
//
// Everything here is a __host__ __device__ function template on <T, Q> that works on ONE cell
// held in registers (T f[Q], indices known at compile time after unrolling).  The expression
// shapes (association order, which literals are T and which are bare doubles) follow the
// reference, because the parity contract is bit-exact populations: the library is compiled with
// -fmad=false and IEEE division / square root.  Citations are relative to /root/reference/src.
#pragma once
#include <cstdint>
#include <cmath>

#define PLB_HD __host__ __device__ __forceinline__

namespace plb200 {

enum DynType : int { DYN_BGK = 0, DYN_RLB = 1, DYN_TRT = 2, DYN_SMAGO = 3, DYN_BOUNCEBACK = 4, DYN_NODYNAMICS = 5, DYN_INCBGK = 6 };
/// not a dynamics: the MAIN argument of the sweep instantiation that collides with externally supplied moments (collide_external)
constexpr int DYN_EXTERNAL = 100;
/// ... and the form that takes them from the fields only where they are not a function of the cell's own populations
constexpr int DYN_EXTERNAL_FUSED = 101;
/// BC_INNER_EDGE / BC_INNER_CORNER: RegularizedVelocityInnerEdgeDynamics3D<plane,normal1,normal2> / ...InnerCornerDynamics3D<xN,yN,zN>
/// (boundaryCondition/regularizedBoundaryDynamics3D.h:40-140).  Record encoding: edge  dir = plane, ori = 2*(normal1 > 0) + (normal2 > 0),
/// with direction1 = (plane+1)%3, direction2 = (plane+2)%3;  corner  ori = 4*(xN > 0) + 2*(yN > 0) + (zN > 0).
enum BcKind : int { BC_VEL_REG = 0, BC_RHO_REG = 1, BC_VEL_ZH = 2, BC_RHO_ZH = 3, BC_STORE_VEL = 4, BC_INNER_EDGE = 5, BC_INNER_CORNER = 6 };
enum ProcClass : int { PROC_PLANE = 0, PROC_EDGE = 1, PROC_CORNER = 2, PROC_PLANE_FD = 3 };

/// One entry of the per-lattice dynamics table (a cell stores the 1-byte index of its entry).
template <class T>
struct Dyn {
    int type;
    T omega;      ///< relaxation parameter (omega0 for Smagorinsky)
    T preFactor;  ///< Smagorinsky: 0.5*(cSmago*omega0*invCs2)^2, complexDynamics/smagorinskyDynamics.hh:45-48
    T rho;        ///< BounceBack / NoDynamics: the density they report, core/dynamics.hh:813-815
};

/// Data of the boundary composite dynamics sitting on a cell (boundaryCondition/boundaryDynamics.hh):
/// stored wall velocity (StoreVelocityDynamics :324-339) or stored rhoBar (StoreDensityDynamics :239-257).
template <class T>
struct Bc {
    int kind, dir, ori;
    T u[3];
    T rhoBar;
};

// ------------------------------------------------------------------------------ descriptors

/// latticeBoltzmann/nearestNeighborLattices3D.hh:165-177 (D3Q19 c), :225-239 (D3Q27 c)
PLB_HD constexpr int c19(int i, int d) {
    constexpr int tab[19][3] = {
        { 0, 0, 0},
        {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1},
        {-1,-1, 0}, {-1, 1, 0}, {-1, 0,-1},
        {-1, 0, 1}, { 0,-1,-1}, { 0,-1, 1},
        { 1, 0, 0}, { 0, 1, 0}, { 0, 0, 1},
        { 1, 1, 0}, { 1,-1, 0}, { 1, 0, 1},
        { 1, 0,-1}, { 0, 1, 1}, { 0, 1,-1} };
    return tab[i][d];
}
PLB_HD constexpr int c27(int i, int d) {
    constexpr int tab[27][3] = {
        { 0, 0, 0},
        {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1},
        {-1,-1, 0}, {-1, 1, 0}, {-1, 0,-1},
        {-1, 0, 1}, { 0,-1,-1}, { 0,-1, 1},
        {-1,-1,-1}, {-1,-1, 1}, {-1, 1,-1}, {-1, 1, 1},
        { 1, 0, 0}, { 0, 1, 0}, { 0, 0, 1},
        { 1, 1, 0}, { 1,-1, 0}, { 1, 0, 1},
        { 1, 0,-1}, { 0, 1, 1}, { 0, 1,-1},
        { 1, 1, 1}, { 1, 1,-1}, { 1,-1, 1}, { 1,-1,-1} };
    return tab[i][d];
}

template <class T, int Q>
struct LB {
    static_assert(Q == 19 || Q == 27, "D3Q19 or D3Q27");
    static constexpr int q = Q;
    static constexpr int H = Q / 2;

    PLB_HD static constexpr int c(int i, int d) { return Q == 19 ? c19(i, d) : c27(i, d); }
    /// indexTemplates::opposite latticeBoltzmann/indexTemplates.h:41-46
    PLB_HD static constexpr int opp(int i) { return i == 0 ? 0 : (i <= H ? i + H : i - H); }
    PLB_HD static constexpr int nn(int i) {
        return (c(i, 0) != 0) + (c(i, 1) != 0) + (c(i, 2) != 0);
    }
    /// weights: nearestNeighborLattices3D.hh:176-196 (D3Q19), :263-277 (D3Q27)
    PLB_HD static constexpr T t(int i) {
        return Q == 19
            ? (nn(i) == 0 ? (T)1 / (T)3 : nn(i) == 1 ? (T)1 / (T)18 : (T)1 / (T)36)
            : (nn(i) == 0 ? (T)8 / (T)27 : nn(i) == 1 ? (T)2 / (T)27 : nn(i) == 2 ? (T)1 / (T)54 : (T)1 / (T)216);
    }
    PLB_HD static constexpr T cs2() { return (T)1 / (T)3; }
    PLB_HD static constexpr T invCs2() { return (T)3; }

    /// latticeBoltzmann/roundOffPolicy.h:74-90 (DefaultRoundOffPolicy)
    PLB_HD static T rhoBarOf(T rho) { return rho - (T)1; }
    PLB_HD static T fullRho(T rhoBar) { return rhoBar + (T)1; }
    PLB_HD static T invRhoOf(T rhoBar) { return (T)1 / (rhoBar + (T)1); }

    // ------------------------------------------------------------------------------ moments

    struct Partial { T xm, x0, xp, ym, yp, zm, zp; };

    /// momentTemplates3D.h:351-373 (D3Q19 partial_rho), :43-72 (D3Q27)
    PLB_HD static Partial partial_rho(const T* f) {
        Partial p;
        if constexpr (Q == 19) {
            p.xm = f[1] + f[4] + f[5] + f[6] + f[7];
            p.xp = f[10] + f[13] + f[14] + f[15] + f[16];
            p.ym = f[2] + f[4] + f[8] + f[9] + f[14];
            p.yp = f[5] + f[11] + f[13] + f[17] + f[18];
            p.zm = f[3] + f[6] + f[8] + f[16] + f[18];
            p.zp = f[7] + f[9] + f[12] + f[15] + f[17];
            p.x0 = f[0] + f[2] + f[3] + f[8] + f[9] + f[11] + f[12] + f[17] + f[18];
        } else {
            p.xm = f[1] + f[4] + f[5] + f[6] + f[7] + f[10] + f[11] + f[12] + f[13];
            p.xp = f[14] + f[17] + f[18] + f[19] + f[20] + f[23] + f[24] + f[25] + f[26];
            p.ym = f[2] + f[4] + f[8] + f[9] + f[10] + f[11] + f[18] + f[25] + f[26];
            p.yp = f[5] + f[12] + f[13] + f[15] + f[17] + f[21] + f[22] + f[23] + f[24];
            p.zm = f[3] + f[6] + f[8] + f[10] + f[12] + f[20] + f[22] + f[24] + f[26];
            p.zp = f[7] + f[9] + f[11] + f[13] + f[16] + f[19] + f[21] + f[23] + f[25];
            p.x0 = f[0] + f[2] + f[3] + f[8] + f[9] + f[15] + f[16] + f[21] + f[22];
        }
        return p;
    }

    /// momentTemplates3D.h:375-382 (D3Q19 get_rhoBar), :74-83 (D3Q27): plain left-to-right sum
    PLB_HD static T get_rhoBar(const T* f) {
        T rhoBar = f[0] + f[1];
#pragma unroll
        for (int i = 2; i < Q; ++i) rhoBar = rhoBar + f[i];
        return rhoBar;
    }

    /// momentTemplates3D.h:384-402 (get_j)
    PLB_HD static void get_j(const T* f, T j[3]) {
        Partial p = partial_rho(f);
        j[0] = (p.xp - p.xm); j[1] = (p.yp - p.ym); j[2] = (p.zp - p.zm);
    }

    /// momentTemplates3D.h:411-423 (D3Q19 get_rhoBar_j), :118-132 (D3Q27)
    PLB_HD static void get_rhoBar_j(const T* f, T& rhoBar, T j[3]) {
        Partial p = partial_rho(f);
        rhoBar = p.xm + p.x0 + p.xp;
        j[0] = (p.xp - p.xm); j[1] = (p.yp - p.ym); j[2] = (p.zp - p.zm);
    }

    /// momentTemplates3D.h:517-539 (D3Q19 compute_rhoBar_j_PiNeq), :239-267 (D3Q27).
    /// Tensor index order xx,xy,xz,yy,yz,zz (geometricOperationTemplates.h:246-248).
    PLB_HD static void rhoBar_j_PiNeq(const T* f, T& rhoBar, T j[3], T PiNeq[6]) {
        Partial p = partial_rho(f);
        rhoBar = p.xm + p.x0 + p.xp;
        j[0] = (p.xp - p.xm); j[1] = (p.yp - p.ym); j[2] = (p.zp - p.zm);
        const T invRho = invRhoOf(rhoBar);
        PiNeq[0] = p.xp + p.xm - cs2() * rhoBar - invRho * j[0] * j[0];
        PiNeq[3] = p.yp + p.ym - cs2() * rhoBar - invRho * j[1] * j[1];
        PiNeq[5] = p.zp + p.zm - cs2() * rhoBar - invRho * j[2] * j[2];
        if constexpr (Q == 19) {
            PiNeq[1] = f[4] - f[5] + f[13] - f[14] - invRho * j[0] * j[1];
            PiNeq[2] = f[6] - f[7] + f[15] - f[16] - invRho * j[0] * j[2];
            PiNeq[4] = f[8] - f[9] + f[17] - f[18] - invRho * j[1] * j[2];
        } else {
            PiNeq[1] = f[4] - f[5] + f[10] + f[11] - f[12] - f[13]
                     + f[17] - f[18] + f[23] + f[24] - f[25] - f[26]
                     - invRho * j[0] * j[1];
            PiNeq[2] = f[6] - f[7] + f[10] - f[11] + f[12] - f[13]
                     + f[19] - f[20] + f[23] - f[24] + f[25] - f[26]
                     - invRho * j[0] * j[2];
            PiNeq[4] = f[8] - f[9] + f[10] - f[11] - f[12] + f[13]
                     + f[21] - f[22] + f[23] - f[24] - f[25] + f[26]
                     - invRho * j[1] * j[2];
        }
    }

    /// momentTemplates::compute_PiNeq(cell, rhoBar, j, PiNeq) latticeBoltzmann/momentTemplates.h:118-123 ->
    /// momentTemplatesImpl::compute_PiNeq(f, rhoBar, j, PiNeq, invRho) momentTemplates3D.h:482-497 (D3Q19), :191-212 (D3Q27):
    /// the off-equilibrium stress from the populations and externally supplied moments
    PLB_HD static void compute_PiNeq(const T* f, T rhoBar, const T j[3], T PiNeq[6]) {
        Partial p = partial_rho(f);
        const T invRho = invRhoOf(rhoBar);
        PiNeq[0] = p.xp + p.xm - cs2() * rhoBar - invRho * j[0] * j[0];
        PiNeq[3] = p.yp + p.ym - cs2() * rhoBar - invRho * j[1] * j[1];
        PiNeq[5] = p.zp + p.zm - cs2() * rhoBar - invRho * j[2] * j[2];
        if constexpr (Q == 19) {
            PiNeq[1] = f[4] - f[5] + f[13] - f[14] - invRho * j[0] * j[1];
            PiNeq[2] = f[6] - f[7] + f[15] - f[16] - invRho * j[0] * j[2];
            PiNeq[4] = f[8] - f[9] + f[17] - f[18] - invRho * j[1] * j[2];
        } else {
            PiNeq[1] = f[4] - f[5] + f[10] + f[11] - f[12] - f[13]
                     + f[17] - f[18] + f[23] + f[24] - f[25] - f[26]
                     - invRho * j[0] * j[1];
            PiNeq[2] = f[6] - f[7] + f[10] - f[11] + f[12] - f[13]
                     + f[19] - f[20] + f[23] - f[24] + f[25] - f[26]
                     - invRho * j[0] * j[2];
            PiNeq[4] = f[8] - f[9] + f[10] - f[11] - f[12] + f[13]
                     + f[21] - f[22] + f[23] - f[24] - f[25] + f[26]
                     - invRho * j[1] * j[2];
        }
    }

    // ------------------------------------------------------------------------------ equilibria

    /// dynamicsTemplates3D.h:1994-1997 (D3Q19: literals are (T)-cast) and :167-170 (D3Q27: bare
    /// double literals, so a float lattice evaluates this in double -- SURVEY.md section 7).
    PLB_HD static T bgk_ma2_equilibrium(int i, T rhoBar, T invRho, const T j[3], T jSqr) {
        T c_j = c(i, 0) * j[0] + c(i, 1) * j[1] + c(i, 2) * j[2];
        if (Q == 19)
            return t(i) * (rhoBar + (T)3. * c_j + invRho * ((T)4.5 * c_j * c_j - (T)1.5 * jSqr));
        return t(i) * (rhoBar + 3. * c_j + invRho * (4.5 * c_j * c_j - 1.5 * jSqr));
    }

    /// dynamicsTemplates3D.h:1999-2076 (D3Q19 bgk_ma2_equilibria), :264-365 (D3Q27)
    PLB_HD static void bgk_ma2_equilibria(T rhoBar, T invRho, const T j[3], T jSqr, T* eq) {
        constexpr int h = H;
        const T t0 = t(0), t1 = t(1), t4 = t(4);
        const T ax = (T)3 * j[0], ay = (T)3 * j[1], az = (T)3 * j[2];
        const T qxx = invRho / (T)2 * ax * ax;
        const T qyy = invRho / (T)2 * ay * ay;
        const T qzz = invRho / (T)2 * az * az;
        const T qxy = invRho * ax * ay;
        const T qxz = invRho * ax * az;
        const T qyz = invRho * ay * az;
        const T iso = rhoBar + invRho * (T)3 * jSqr;
        T lin, quad;
        quad = -qxx - qyy - qzz;           eq[0] = t0 * (iso + quad);
        lin = -ax;       quad = -qyy - qzz;    eq[1] = t1 * (iso + lin + quad); eq[1 + h] = t1 * (iso - lin + quad);
        lin = -ay;       quad = -qxx - qzz;    eq[2] = t1 * (iso + lin + quad); eq[2 + h] = t1 * (iso - lin + quad);
        lin = -az;       quad = -qxx - qyy;    eq[3] = t1 * (iso + lin + quad); eq[3 + h] = t1 * (iso - lin + quad);
        lin = -ax - ay;  quad = qxy - qzz;      eq[4] = t4 * (iso + lin + quad); eq[4 + h] = t4 * (iso - lin + quad);
        lin = -ax + ay;  quad = -qxy - qzz;     eq[5] = t4 * (iso + lin + quad); eq[5 + h] = t4 * (iso - lin + quad);
        lin = -ax - az;  quad = qxz - qyy;      eq[6] = t4 * (iso + lin + quad); eq[6 + h] = t4 * (iso - lin + quad);
        lin = -ax + az;  quad = -qxz - qyy;     eq[7] = t4 * (iso + lin + quad); eq[7 + h] = t4 * (iso - lin + quad);
        lin = -ay - az;  quad = qyz - qxx;      eq[8] = t4 * (iso + lin + quad); eq[8 + h] = t4 * (iso - lin + quad);
        lin = -ay + az;  quad = -qyz - qxx;     eq[9] = t4 * (iso + lin + quad); eq[9 + h] = t4 * (iso - lin + quad);
        if constexpr (Q == 27) {
            const T t10 = t(10);
            lin = -ax - ay - az;  quad = qxy + qxz + qyz;   eq[10] = t10 * (iso + lin + quad); eq[23] = t10 * (iso - lin + quad);
            lin = -ax - ay + az;  quad = qxy - qxz - qyz;   eq[11] = t10 * (iso + lin + quad); eq[24] = t10 * (iso - lin + quad);
            lin = -ax + ay - az;  quad = -qxy + qxz - qyz;  eq[12] = t10 * (iso + lin + quad); eq[25] = t10 * (iso - lin + quad);
            lin = -ax + ay + az;  quad = -qxy - qxz + qyz;  eq[13] = t10 * (iso + lin + quad); eq[26] = t10 * (iso - lin + quad);
        }
    }

    // ------------------------------------------------------------------------------ collisions

    /// D3Q19: dynamicsTemplates3D.h:2289-2368 (bgk_ma2_collision_base) via :2476-2478.
    /// D3Q27: generic loop :1624-1634.  Returns uSqr.
    PLB_HD static T bgk_ma2_collision(T* f, T rhoBar, const T j[3], T omega) {
        const T invRho = invRhoOf(rhoBar);
        if constexpr (Q == 19) {
            // the reference's D3Q19 specialisation IS the explicit-invRho form called with 1/rho (dynamicsTemplates3D.h:2476-2478)
            return bgk_ma2_collision_base(f, rhoBar, j, omega, invRho);
        } else {
            const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                f[i] *= (T)1 - omega;
                f[i] += omega * bgk_ma2_equilibrium(i, rhoBar, invRho, j, jSqr);
            }
            return jSqr * invRho * invRho;
        }
    }

    /// bgk_ma2_collision_base(f, rhoBar, j, omega, invRho) dynamicsTemplates3D.h:2289-2368 (D3Q19), :467-571 (D3Q27): the form
    /// with an explicit invRho, used by bgk_inc_collision (:2480-2482, :1680-1682) with invRho0.  Returns invRho^2 * jSqr.
    PLB_HD static T bgk_ma2_collision_base(T* f, T rhoBar, const T j[3], T omega, T invRho) {
        constexpr int h = H;
        const T keep = (T)1 - omega;
        const T w0 = t(0) * omega;
        const T w1 = t(1) * omega;
        const T w4 = t(4) * omega;
        const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
        const T ax = (T)3 * j[0];
        const T ay = (T)3 * j[1];
        const T az = (T)3 * j[2];
        const T qxx = invRho / (T)2 * ax * ax;
        const T qyy = invRho / (T)2 * ay * ay;
        const T qzz = invRho / (T)2 * az * az;
        const T qxy = invRho * ax * ay;
        const T qxz = invRho * ax * az;
        const T qyz = invRho * ay * az;
        const T iso = rhoBar + invRho * (T)3 * jSqr;
        T lin, quad;
        quad = -qxx - qyy - qzz;
        f[0] *= keep; f[0] += w0 * (iso + quad);
#define PLB_PAIR(I, TW) f[I] *= keep; f[I] += TW * (iso + lin + quad); f[I + h] *= keep; f[I + h] += TW * (iso - lin + quad);
        lin = -ax;       quad = -qyy - qzz;   PLB_PAIR(1, w1)
        lin = -ay;       quad = -qxx - qzz;   PLB_PAIR(2, w1)
        lin = -az;       quad = -qxx - qyy;   PLB_PAIR(3, w1)
        lin = -ax - ay;  quad = qxy - qzz;     PLB_PAIR(4, w4)
        lin = -ax + ay;  quad = -qxy - qzz;    PLB_PAIR(5, w4)
        lin = -ax - az;  quad = qxz - qyy;     PLB_PAIR(6, w4)
        lin = -ax + az;  quad = -qxz - qyy;    PLB_PAIR(7, w4)
        lin = -ay - az;  quad = qyz - qxx;     PLB_PAIR(8, w4)
        lin = -ay + az;  quad = -qyz - qxx;    PLB_PAIR(9, w4)
        if constexpr (Q == 27) {
            const T w10 = t(10) * omega;
            lin = -ax - ay - az;  quad = qxy + qxz + qyz;    PLB_PAIR(10, w10)
            lin = -ax - ay + az;  quad = qxy - qxz - qyz;    PLB_PAIR(11, w10)
            lin = -ax + ay - az;  quad = -qxy + qxz - qyz;   PLB_PAIR(12, w10)
            lin = -ax + ay + az;  quad = -qxy - qxz + qyz;   PLB_PAIR(13, w10)
        }
#undef PLB_PAIR
        return invRho * invRho * jSqr;
    }

    /// neqPiD3Q19 dynamicsTemplates3D.h:1903-1975 and neqPiD3Q27 :38-135 (fromPiToFneq0..9 / 0..13)
    PLB_HD static void neqPi(const T* pi, T* n) {
        const T xx = pi[0], xy = pi[1], xz = pi[2], yy = pi[3], yz = pi[4], zz = pi[5];
        if constexpr (Q == 19) {
            n[0] = (T)3. / (T)2 * (-(T)1. / (T)3 * xx - (T)1. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[1] = (T)1. / (T)4 * ((T)2. / (T)3 * xx - (T)1. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[2] = (T)1. / (T)4 * (-(T)1. / (T)3 * xx + (T)2. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[3] = (T)1. / (T)4 * (-(T)1. / (T)3 * xx - (T)1. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[4] = (T)1. / (T)8 * ((T)2. / (T)3 * xx + (T)2. / (T)3 * yy - (T)1. / (T)3 * zz + (T)2 * xy);
            n[5] = (T)1. / (T)8 * ((T)2. / (T)3 * xx + (T)2. / (T)3 * yy - (T)1. / (T)3 * zz - (T)2 * xy);
            n[6] = (T)1. / (T)8 * ((T)2. / (T)3 * xx - (T)1. / (T)3 * yy + (T)2. / (T)3 * zz + (T)2 * xz);
            n[7] = (T)1. / (T)8 * ((T)2. / (T)3 * xx - (T)1. / (T)3 * yy + (T)2. / (T)3 * zz - (T)2 * xz);
            n[8] = (T)1. / (T)8 * (-(T)1. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz + (T)2 * yz);
            n[9] = (T)1. / (T)8 * (-(T)1. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz - (T)2 * yz);
        } else {
            n[0] = (T)4. / (T)3 * (-(T)1. / (T)3 * xx - (T)1. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[1] = (T)1. / (T)3 * ((T)2. / (T)3 * xx - (T)1. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[2] = (T)1. / (T)3 * (-(T)1. / (T)3 * xx + (T)2. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[3] = (T)1. / (T)3 * (-(T)1. / (T)3 * xx - (T)1. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[4] = (T)1. / (T)12 * ((T)2 * xy + (T)2. / (T)3 * xx + (T)2. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[5] = (T)1. / (T)12 * (-(T)2 * xy + (T)2. / (T)3 * xx + (T)2. / (T)3 * yy - (T)1. / (T)3 * zz);
            n[6] = (T)1. / (T)12 * ((T)2 * xz + (T)2. / (T)3 * xx - (T)1. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[7] = (T)1. / (T)12 * (-(T)2 * xz + (T)2. / (T)3 * xx - (T)1. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[8] = (T)1. / (T)12 * ((T)2 * yz - (T)1. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[9] = (T)1. / (T)12 * (-(T)2 * yz - (T)1. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[10] = (T)1. / (T)48 * (+(T)2 * xy + (T)2 * xz + (T)2 * yz
                                     + (T)2. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[11] = (T)1. / (T)48 * (+(T)2 * xy - (T)2 * xz - (T)2 * yz
                                     + (T)2. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[12] = (T)1. / (T)48 * (-(T)2 * xy + (T)2 * xz - (T)2 * yz
                                     + (T)2. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz);
            n[13] = (T)1. / (T)48 * (-(T)2 * xy - (T)2 * xz + (T)2 * yz
                                     + (T)2. / (T)3 * xx + (T)2. / (T)3 * yy + (T)2. / (T)3 * zz);
        }
    }

    /// dynamicsTemplates3D.h:2500-2554 (D3Q19 rlb_collision), :1683-1760 (D3Q27)
    PLB_HD static T rlb_collision(T* f, T rhoBar, T invRho, const T j[3], const T PiNeq[6], T omega) {
        const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
        T n[14];
        neqPi(PiNeq, n);
        f[0] = bgk_ma2_equilibrium(0, rhoBar, invRho, j, jSqr) + ((T)1 - omega) * n[0];
#pragma unroll
        for (int i = 1; i <= H; ++i)
            f[i] = bgk_ma2_equilibrium(i, rhoBar, invRho, j, jSqr) + ((T)1 - omega) * n[i];
#pragma unroll
        for (int i = 1; i <= H; ++i)
            f[i + H] = bgk_ma2_equilibrium(i + H, rhoBar, invRho, j, jSqr) + ((T)1 - omega) * n[i];
        return jSqr * invRho * invRho;
    }

    /// complexDynamics/smagorinskyDynamics.hh:45-48 (computePrefactor; util::sqr(x) = x*x)
    PLB_HD static T smago_prefactor(T omega0, T cSmago) {
        T x = cSmago * omega0 * invCs2();
        return (T)0.5 * (x * x);
    }
    /// smagorinskyDynamics.hh:50-59 (computeOmega) + tensorNormSqr geometricOperationTemplates.h:254-257
    PLB_HD static T smago_omega(T omega0, T preFactor, T rhoBar, const T m[6]) {
        T PiNeqNormSqr = m[0] * m[0] + m[3] * m[3] + m[5] * m[5]
                       + (T)2 * m[1] * m[1] + (T)2 * m[2] * m[2] + (T)2 * m[4] * m[4];
        T PiNeqNorm = sqrt(PiNeqNormSqr);
        T alpha = preFactor * invRhoOf(rhoBar);
        T linearTerm = alpha * PiNeqNorm;
        T squareTerm = (T)2 * alpha * alpha * PiNeqNormSqr;
        return omega0 * (1 - linearTerm + squareTerm);
    }

    /// BGKdynamics::collide basicDynamics/isoThermalDynamics.hh:319-331
    PLB_HD static void collide_bgk(T* f, T omega, T& rhoBar, T& uSqr) {
        T j[3];
        get_rhoBar_j(f, rhoBar, j);
        uSqr = bgk_ma2_collision(f, rhoBar, j, omega);
    }
    /// IncBGKdynamics::collide isoThermalDynamics.hh:1405-1417: BGK whose O(Ma^2) terms carry 1/rho0 (= 1) instead of 1/rho
    PLB_HD static void collide_incbgk(T* f, T omega, T& rhoBar, T& uSqr) {
        T j[3];
        get_rhoBar_j(f, rhoBar, j);
        uSqr = bgk_ma2_collision_base(f, rhoBar, j, omega, (T)1);
    }
    /// RegularizedBGKdynamics::collide isoThermalDynamics.hh:1588-1602
    PLB_HD static void collide_rlb(T* f, T omega, T& rhoBar, T& uSqr) {
        T j[3], PiNeq[6];
        rhoBar_j_PiNeq(f, rhoBar, j, PiNeq);
        const T invRho = invRhoOf(rhoBar);
        uSqr = rlb_collision(f, rhoBar, invRho, j, PiNeq, omega);
    }
    /// TRTdynamics::collide complexDynamics/trtDynamics.hh:64-99, sMinus = 1.1 (:40)
    PLB_HD static void collide_trt(T* f, T omega, T& rhoBar, T& uSqr) {
        const T sPlus = omega;
        const T sMinus = (T)1.1;
        T j[3], eq[Q];
        get_rhoBar_j(f, rhoBar, j);
        const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
        const T invRho = invRhoOf(rhoBar);
        bgk_ma2_equilibria(rhoBar, invRho, j, jSqr, eq);
        f[0] += -sPlus * f[0] + sPlus * eq[0];
#pragma unroll
        for (int i = 1; i <= H; ++i) {
            const T eq_plus = 0.5 * (eq[i] + eq[i + H]);
            const T eq_minus = 0.5 * (eq[i] - eq[i + H]);
            const T f_plus = 0.5 * (f[i] + f[i + H]);
            const T f_minus = 0.5 * (f[i] - f[i + H]);
            f[i]     += -sPlus * (f_plus - eq_plus) - sMinus * (f_minus - eq_minus);
            f[i + H] += -sPlus * (f_plus - eq_plus) + sMinus * (f_minus - eq_minus);
        }
        uSqr = jSqr * invRho * invRho;
    }
    /// SmagorinskyBGKdynamics::collide complexDynamics/smagorinskyDynamics.hh:290-304
    PLB_HD static void collide_smago(T* f, T omega0, T preFactor, T& rhoBar, T& uSqr) {
        T j[3], PiNeq[6];
        rhoBar_j_PiNeq(f, rhoBar, j, PiNeq);
        const T omega = smago_omega(omega0, preFactor, rhoBar, PiNeq);
        uSqr = bgk_ma2_collision(f, rhoBar, j, omega);
    }
    /// BounceBack::collide core/dynamics.hh:783-791
    PLB_HD static void collide_bounceback(T* f) {
#pragma unroll
        for (int i = 1; i <= H; ++i) { T tmp = f[i]; f[i] = f[i + H]; f[i + H] = tmp; }
    }

    /// Cell::collide -> Dynamics::collide (core/cell.h:183-186) with the dynamics type known at
    /// compile time.  Returns true (and rhoBar, uSqr) when the dynamics gathers statistics.
    template <int TYPE>
    PLB_HD static bool collide_as(T* f, T omega, T preFactor, T& rhoBar, T& uSqr) {
        if (TYPE == DYN_BGK) { collide_bgk(f, omega, rhoBar, uSqr); return true; }
        if (TYPE == DYN_RLB) { collide_rlb(f, omega, rhoBar, uSqr); return true; }
        if (TYPE == DYN_TRT) { collide_trt(f, omega, rhoBar, uSqr); return true; }
        if (TYPE == DYN_SMAGO) { collide_smago(f, omega, preFactor, rhoBar, uSqr); return true; }
        if (TYPE == DYN_INCBGK) { collide_incbgk(f, omega, rhoBar, uSqr); return true; }
        if (TYPE == DYN_BOUNCEBACK) { collide_bounceback(f); return false; }
        return false;   // NoDynamics core/dynamics.hh:968-972
    }

    /// Same with the type read from the dynamics table at run time.
    PLB_HD static bool collide_any(const Dyn<T>& d, T* f, T& rhoBar, T& uSqr) {
        switch (d.type) {
        case DYN_BGK:        collide_bgk(f, d.omega, rhoBar, uSqr); return true;
        case DYN_RLB:        collide_rlb(f, d.omega, rhoBar, uSqr); return true;
        case DYN_TRT:        collide_trt(f, d.omega, rhoBar, uSqr); return true;
        case DYN_SMAGO:      collide_smago(f, d.omega, d.preFactor, rhoBar, uSqr); return true;
        case DYN_INCBGK:     collide_incbgk(f, d.omega, rhoBar, uSqr); return true;
        case DYN_BOUNCEBACK: collide_bounceback(f); return false;
        default:             return false;
        }
    }

    /// Dynamics::collideExternal(cell, rhoBar, j, thetaBar, stat): the collision with moments taken from fields instead of
    /// the populations (ExternalRhoJcollideAndStream3D, basicDynamics/dynamicsProcessor3D.hh:41-76).  BGK
    /// isoThermalDynamics.hh:333-338, IncBGK :1422-1427, RegularizedBGK :1605-1615, TRT complexDynamics/trtDynamics.hh:102-134,
    /// SmagorinskyBGK smagorinskyDynamics.hh:306-320, BounceBack core/dynamics.hh:793-799, NoDynamics :974-978; a boundary
    /// composite forwards to its base dynamics (core/dynamics.hh:478-484).
    PLB_HD static void collide_external(const Dyn<T>& d, T* f, T rhoBar, const T j[3]) {
        switch (d.type) {
        case DYN_BGK:    bgk_ma2_collision(f, rhoBar, j, d.omega); break;
        case DYN_INCBGK: bgk_ma2_collision_base(f, rhoBar, j, d.omega, (T)1); break;
        case DYN_RLB: {
            T PiNeq[6];
            compute_PiNeq(f, rhoBar, j, PiNeq);
            const T invRho = invRhoOf(rhoBar);
            rlb_collision(f, rhoBar, invRho, j, PiNeq, d.omega);
            break; }
        case DYN_TRT: {
            const T sPlus = d.omega;
            const T sMinus = (T)1.1;
            T eq[Q];
            const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
            const T invRho = invRhoOf(rhoBar);
            bgk_ma2_equilibria(rhoBar, invRho, j, jSqr, eq);
            f[0] += -sPlus * f[0] + sPlus * eq[0];
#pragma unroll
            for (int i = 1; i <= H; ++i) {
                const T eq_plus = 0.5 * (eq[i] + eq[i + H]);
                const T eq_minus = 0.5 * (eq[i] - eq[i + H]);
                const T f_plus = 0.5 * (f[i] + f[i + H]);
                const T f_minus = 0.5 * (f[i] - f[i + H]);
                f[i]     += -sPlus * (f_plus - eq_plus) - sMinus * (f_minus - eq_minus);
                f[i + H] += -sPlus * (f_plus - eq_plus) + sMinus * (f_minus - eq_minus);
            }
            break; }
        case DYN_SMAGO: {
            T PiNeq[6];
            compute_PiNeq(f, rhoBar, j, PiNeq);
            const T omega = smago_omega(d.omega, d.preFactor, rhoBar, PiNeq);
            bgk_ma2_collision(f, rhoBar, j, omega);
            break; }
        case DYN_BOUNCEBACK: collide_bounceback(f); break;
        default: break;
        }
    }

    // ------------------------------------------------------------------------ regularization

    PLB_HD static bool is_inert(int type) { return type == DYN_BOUNCEBACK || type == DYN_NODYNAMICS; }

    /// Dynamics::computeEquilibrium(iPop, rhoBar, j, jSqr): identical for BGK / RLB / TRT /
    /// Smagorinsky (isoThermalDynamics.hh:341-346, :1619-1624, trtDynamics.hh:138-143,
    /// smagorinskyDynamics.hh:322-328); BounceBack / NoDynamics return 0 (core/dynamics.hh:800-805); IncBGK uses
    /// invRho0 = 1 (isoThermalDynamics.hh:1428-1435).
    /// INC: the base dynamics is IncBGK -- a compile-time fact in the plane-completion family below, so that the common
    /// (compressible) instantiation carries no trace of it
    template <bool INC>
    PLB_HD static T computeEquilibriumT(int type, int i, T rhoBar, const T j[3], T jSqr) {
        if (is_inert(type)) return (T)0;
        const T invRho = INC ? (T)1 : invRhoOf(rhoBar);
        return bgk_ma2_equilibrium(i, rhoBar, invRho, j, jSqr);
    }
    PLB_HD static T computeEquilibrium(int type, int i, T rhoBar, const T j[3], T jSqr) {
        return type == DYN_INCBGK ? computeEquilibriumT<true>(type, i, rhoBar, j, jSqr) : computeEquilibriumT<false>(type, i, rhoBar, j, jSqr);
    }

    /// latticeBoltzmann/offEquilibriumTemplates.h:91-114 (generic; no 3D specialisation exists)
    PLB_HD static T fromPiToFneq(int i, const T PiNeq[6]) {
        T fNeq = (T)0;
        int iPi = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            fNeq += PiNeq[iPi] * (c(i, a) * c(i, a) - cs2());
            ++iPi;
#pragma unroll
            for (int b = a + 1; b < 3; ++b) {
                fNeq += PiNeq[iPi] * (T)2 * c(i, a) * c(i, b);
                ++iPi;
            }
        }
        fNeq *= t(i) * invCs2() * invCs2() / (T)2;
        return fNeq;
    }

    /// IsoThermalBulkDynamics::regularize basicDynamics/isoThermalDynamics.hh:54-70;
    /// BounceBack/NoDynamics::regularize are no-ops (core/dynamics.hh:807-811, :986-990)
    template <bool INC>
    PLB_HD static void regularizeT(int type, T* f, T rhoBar, const T j[3], T jSqr, const T PiNeq[6]) {
        if (is_inert(type)) return;
        f[0] = computeEquilibriumT<INC>(type, 0, rhoBar, j, jSqr) + fromPiToFneq(0, PiNeq);
#pragma unroll
        for (int i = 1; i <= H; ++i) {
            f[i] = computeEquilibriumT<INC>(type, i, rhoBar, j, jSqr);
            f[i + H] = computeEquilibriumT<INC>(type, i + H, rhoBar, j, jSqr);
            const T fNeq = fromPiToFneq(i, PiNeq);
            f[i] += fNeq;
            f[i + H] += fNeq;
        }
    }
    PLB_HD static void regularize(int type, T* f, T rhoBar, const T j[3], T jSqr, const T PiNeq[6]) {
        if (type == DYN_INCBGK) regularizeT<true>(type, f, rhoBar, j, jSqr, PiNeq);
        else regularizeT<false>(type, f, rhoBar, j, jSqr, PiNeq);
    }

    // ---------------------------------------------------------------- plane boundary dynamics

    PLB_HD static int cdir(int i, int dir) { return dir == 0 ? c(i, 0) : (dir == 1 ? c(i, 1) : c(i, 2)); }

    /// sums over indexTemplates::subIndex<D,direction,value>() in ascending index order
    /// (latticeBoltzmann/indexTemplates.h:78-99)
    PLB_HD static T sub_sum(const T* f, int dir, int value) {
        T s = (T)0;
#pragma unroll
        for (int i = 0; i < Q; ++i) if (cdir(i, dir) == value) s += f[i];
        return s;
    }

    /// VelocityDirichletBoundaryDynamics::computeRhoBar boundaryCondition/boundaryDynamics.hh:639-668
    /// (the descriptors on this path carry no external force: the 0.5*force terms vanish)
    /// velIsJ: the base dynamics treats the velocity as momentum (Dynamics::velIsJ, true for IncBGKdynamics,
    /// isoThermalDynamics.hh:1437-1440; composites forward it, core/dynamics.hh:424-426)
    PLB_HD static bool vel_is_j(int baseType) { return baseType == DYN_INCBGK; }

    PLB_HD static T veldir_rhoBar(const Bc<T>& b, const T* f, bool velIsJ) {
        const T rhoOnWall = sub_sum(f, b.dir, 0);
        const T rhoNormal = sub_sum(f, b.dir, b.ori);
        const T uN = b.dir == 0 ? b.u[0] : (b.dir == 1 ? b.u[1] : b.u[2]);
        const T velNormal = (T)b.ori * (uN - 0.5 * (T)0);
        if (velIsJ) return (T)2 * rhoNormal + rhoOnWall - velNormal;
        return ((T)2 * rhoNormal + rhoOnWall - 1 * velNormal) / ((T)1 + velNormal);
    }

    /// DensityDirichletBoundaryDynamics::computeJ boundaryDynamics.hh:760-787
    PLB_HD static void rhodir_j(const Bc<T>& b, const T* f, T j[3]) {
        const T rhoOnWall = sub_sum(f, b.dir, 0);
        const T rhoNormal = sub_sum(f, b.dir, b.ori);
        const T jn = (T)b.ori * ((T)2 * rhoNormal + rhoOnWall - b.rhoBar);
        j[0] = b.dir == 0 ? jn : (T)0;
        j[1] = b.dir == 1 ? jn : (T)0;
        j[2] = b.dir == 2 ? jn : (T)0;
    }

    PLB_HD static bool is_vel_plane(int kind) { return kind == BC_VEL_REG || kind == BC_VEL_ZH; }
    PLB_HD static bool is_rho_plane(int kind) { return kind == BC_RHO_REG || kind == BC_RHO_ZH; }

    /// computeRhoBarJ of the plane boundary dynamics: velocity boundaryDynamics.hh:680-701, density :822-828
    PLB_HD static void plane_rhoBar_j(const Bc<T>& b, const T* f, T& rhoBar, T j[3], bool velIsJ) {
        if (is_vel_plane(b.kind)) {
            rhoBar = veldir_rhoBar(b, f, velIsJ);
            const T rho = fullRho(rhoBar);
#pragma unroll
            for (int d = 0; d < 3; ++d) j[d] = velIsJ ? (b.u[d] - 0.5 * (T)0) : rho * (b.u[d] - 0.5 * (T)0);
        } else {
            rhoBar = b.rhoBar;
            rhodir_j(b, f, j);
        }
    }

    // ------------------------------------------------- concave (internal) edges and corners

    PLB_HD static bool is_inner(int kind) { return kind == BC_INNER_EDGE || kind == BC_INNER_CORNER; }
    /// the velocity-Dirichlet plane dynamics an inner edge (two) or corner (three) is composed of; all share the stored velocity
    PLB_HD static int inner_parts(const Bc<T>& b, Bc<T> part[3]) {
        if (b.kind == BC_INNER_EDGE) {
            part[0] = b; part[0].kind = BC_VEL_REG; part[0].dir = (b.dir + 1) % 3; part[0].ori = (b.ori & 2) ? 1 : -1;
            part[1] = b; part[1].kind = BC_VEL_REG; part[1].dir = (b.dir + 2) % 3; part[1].ori = (b.ori & 1) ? 1 : -1;
            part[2] = part[1];
            return 2;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { part[k] = b; part[k].kind = BC_VEL_REG; part[k].dir = k; part[k].ori = (b.ori & (4 >> k)) ? 1 : -1; }
        return 3;
    }
    /// computeRhoBar regularizedBoundaryDynamics3D.hh:95-101 (edge), :237-245 (corner)
    PLB_HD static T inner_rhoBar(const Bc<T>& b, const T* f, bool velIsJ) {
        Bc<T> part[3];
        if (inner_parts(b, part) == 2) return (veldir_rhoBar(part[0], f, velIsJ) + veldir_rhoBar(part[1], f, velIsJ)) / (T)2;
        return (veldir_rhoBar(part[0], f, velIsJ) + veldir_rhoBar(part[1], f, velIsJ) + veldir_rhoBar(part[2], f, velIsJ)) / (T)3;
    }
    /// computeDensity :86-92, :225-234: the mean of the parts' densities
    PLB_HD static T inner_density(const Bc<T>& b, const T* f, bool velIsJ) {
        Bc<T> part[3];
        if (inner_parts(b, part) == 2) return (fullRho(veldir_rhoBar(part[0], f, velIsJ)) + fullRho(veldir_rhoBar(part[1], f, velIsJ))) / (T)2;
        return (fullRho(veldir_rhoBar(part[0], f, velIsJ)) + fullRho(veldir_rhoBar(part[1], f, velIsJ)) + fullRho(veldir_rhoBar(part[2], f, velIsJ))) / (T)3;
    }
    /// computeRhoBarJ :104-116, :248-260: j = stored velocity, times rho unless velIsJ
    PLB_HD static void inner_rhoBar_j(const Bc<T>& b, const T* f, T& rhoBar, T j[3], bool velIsJ) {
        rhoBar = inner_rhoBar(b, f, velIsJ);
        const T rho = fullRho(rhoBar);
#pragma unroll
        for (int d = 0; d < 3; ++d) { j[d] = b.u[d]; if (!velIsJ) j[d] *= rho; }
    }
    /// completePopulations :138-162 (edge), :276-298 (corner): bounce-back of off-equilibrium on the populations that leave the solid
    /// along both (all three) normals, then regularize ALL populations with the base dynamics' PiNeq of the cell
    /// (IsoThermalBulkDynamics::computePiNeq isoThermalDynamics.hh:78-85) around the boundary's own rhoBar and j
    template <bool INC>
    PLB_HD static void complete_inner(int baseType, const Bc<T>& b, T* f) {
        Bc<T> part[3];
        const int n = inner_parts(b, part);
        T rhoBar, j[3], PiNeq[6];
        inner_rhoBar_j(b, f, rhoBar, j, INC);
        const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            bool unknown = cdir(i, part[0].dir) == -part[0].ori && cdir(i, part[1].dir) == -part[1].ori;
            if (n == 3) unknown = unknown && cdir(i, part[2].dir) == -part[2].ori;
            if (unknown)
                f[i] = f[opp(i)] - computeEquilibriumT<INC>(baseType, opp(i), rhoBar, j, jSqr)
                                 + computeEquilibriumT<INC>(baseType, i, rhoBar, j, jSqr);
        }
        if (is_inert(baseType)) {
#pragma unroll
            for (int k = 0; k < 6; ++k) PiNeq[k] = (T)0;           // core/dynamics.hh:837-842, :1017-1022
        } else {
            T cellRhoBar, cellJ[3];
            get_rhoBar_j(f, cellRhoBar, cellJ);
            compute_PiNeq(f, cellRhoBar, cellJ, PiNeq);
        }
        regularizeT<INC>(baseType, f, rhoBar, j, jSqr, PiNeq);
    }

    /// boundaryTemplates<..>::compute_PiNeq boundaryCondition/boundaryTemplates.h:45-87
    template <bool INC>
    PLB_HD static void boundary_PiNeq(int baseType, const Bc<T>& b, const T* f, T rhoBar, const T j[3], T jSqr, T PiNeq[6]) {
        T fNeq[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            const int cd = cdir(i, b.dir);
            if (cd == 0) fNeq[i] = f[i] - computeEquilibriumT<INC>(baseType, i, rhoBar, j, jSqr);
            else if (cd == b.ori) fNeq[i] = f[i] - computeEquilibriumT<INC>(baseType, i, rhoBar, j, jSqr);
            else fNeq[i] = (T)0;
        }
        int iPi = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int bb = a; bb < 3; ++bb) {
                T s = (T)0;
#pragma unroll
                for (int i = 0; i < Q; ++i) if (cdir(i, b.dir) == 0)
                    s += c(i, a) * c(i, bb) * fNeq[i];
#pragma unroll
                for (int i = 0; i < Q; ++i) if (cdir(i, b.dir) == b.ori)
                    s += (T)2 * c(i, a) * c(i, bb) * fNeq[i];
                PiNeq[iPi] = s;
                ++iPi;
            }
        }
    }

    /// ZouHeClosure boundaryCondition/zouHeDynamics.hh:43-115
    template <bool INC>
    PLB_HD static void zouhe_closure(int baseType, const Bc<T>& b, T* f) {
        // missing populations = opposites of subIndex<dir,ori>; the diagonal ones (all but the single
        // axis-aligned one) share the transverse-momentum correction: 4 for D3Q19, 8 for D3Q27.
        constexpr int nDiag = (Q == 19) ? 4 : 8;
        T rhoBar, j[3];
        plane_rhoBar_j(b, f, rhoBar, j, INC);
        const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
#pragma unroll
        for (int i = 1; i < Q; ++i) if (cdir(i, b.dir) == b.ori) {
            // pop = opp(i) is missing, its opposite i is known
            f[opp(i)] = f[i] - computeEquilibriumT<INC>(baseType, i, rhoBar, j, jSqr)
                             + computeEquilibriumT<INC>(baseType, opp(i), rhoBar, j, jSqr);
        }
        // momentTemplates::compute_rho_uLb momentTemplates3D.h:437-445
        T falseRhoBar, falseJ[3], falseU[3], diff[3];
        get_rhoBar_j(f, falseRhoBar, falseJ);
        const T invRho = invRhoOf(falseRhoBar);
        const T falseRho = fullRho(falseRhoBar);
#pragma unroll
        for (int d = 0; d < 3; ++d) { falseU[d] = falseJ[d]; falseU[d] *= invRho; }
#pragma unroll
        for (int d = 0; d < 3; ++d) diff[d] = (j[d] - falseRho * falseU[d]) / (T)nDiag;
        const int d1 = (b.dir + 1) % 3, d2 = (b.dir + 2) % 3;
        const T diff1 = d1 == 0 ? diff[0] : (d1 == 1 ? diff[1] : diff[2]);
        const T diff2 = d2 == 0 ? diff[0] : (d2 == 1 ? diff[1] : diff[2]);
#pragma unroll
        for (int i = 1; i < Q; ++i) if (cdir(i, b.dir) == b.ori && nn(i) > 1) {
            const int p = opp(i);
            f[p] += cdir(p, d1) * diff1;
            f[p] += cdir(p, d2) * diff2;
        }
    }

    /// completePopulations of the four plane boundary dynamics:
    /// RegularizedVelocity/DensityBoundaryDynamics regularizedBoundaryDynamics.hh:94-109, :166-181;
    /// ZouHeVelocity/PressureDynamics zouHeDynamics.hh:166-169, :222-225.  StoreVelocityDynamics
    /// (edges, corners) inherits the empty completePopulations (boundaryDynamics.hh:185-187).
    template <bool INC>
    PLB_HD static void complete_populations(int baseType, const Bc<T>& b, T* f) {
        if (b.kind == BC_VEL_REG || b.kind == BC_RHO_REG) {
            T rhoBar, j[3], PiNeq[6];
            plane_rhoBar_j(b, f, rhoBar, j, INC);
            const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
            boundary_PiNeq<INC>(baseType, b, f, rhoBar, j, jSqr, PiNeq);
            regularizeT<INC>(baseType, f, rhoBar, j, jSqr, PiNeq);
        } else if (b.kind == BC_VEL_ZH || b.kind == BC_RHO_ZH) {
            zouhe_closure<INC>(baseType, b, f);
        } else if (is_inner(b.kind)) {
            complete_inner<INC>(baseType, b, f);
        }
    }

    // --------------------------------------------- Cell::computeDensity / Velocity / RhoBar

    /// Dynamics::computeDensity of whatever sits on the cell (hasBc: a boundary composite is on top):
    ///  bulk BasicBulkDynamics::computeDensity core/dynamics.hh:277-280; BB/NoDynamics their rho
    ///  (:813-815, :992-994); velocity planes boundaryDynamics.hh:671-677; density planes :247-250;
    ///  edges/corners fall through to the base dynamics (:83-87).
    PLB_HD static T cell_density(const Dyn<T>& d, bool hasBc, const Bc<T>& b, const T* f) {
        if (hasBc) {
            if (is_vel_plane(b.kind)) return fullRho(veldir_rhoBar(b, f, vel_is_j(d.type)));
            if (is_rho_plane(b.kind)) return fullRho(b.rhoBar);
            if (is_inner(b.kind)) return inner_density(b, f, vel_is_j(d.type));
        }
        if (is_inert(d.type)) return d.rho;
        return fullRho(get_rhoBar(f));
    }
    /// True when cell_density needs the populations of the cell.
    PLB_HD static bool density_reads_pops(int dynType, bool hasBc, int bcKind) {
        if (hasBc) {
            if (is_vel_plane(bcKind) || is_inner(bcKind)) return true;
            if (is_rho_plane(bcKind)) return false;
        }
        return !is_inert(dynType);
    }

    /// Dynamics::computeVelocity: bulk compute_uLb momentTemplates3D.h:429-435; BB/NoDynamics zero
    /// (core/dynamics.hh:821-826); StoreVelocityDynamics (velocity planes, edges, corners)
    /// boundaryDynamics.hh:324-331; density planes :790-812.
    PLB_HD static void cell_velocity(const Dyn<T>& d, bool hasBc, const Bc<T>& b, const T* f, T u[3]) {
        if (hasBc) {
            if (is_rho_plane(b.kind)) {
                const T invRho = invRhoOf(b.rhoBar);
                rhodir_j(b, f, u);
#pragma unroll
                for (int k = 0; k < 3; ++k) { if (!vel_is_j(d.type)) u[k] *= invRho; u[k] += 0.5 * (T)0; }
                return;
            }
            u[0] = b.u[0]; u[1] = b.u[1]; u[2] = b.u[2];
            return;
        }
        if (is_inert(d.type)) { u[0] = u[1] = u[2] = (T)0; return; }
        if (d.type == DYN_INCBGK) {            // IncBGKdynamics::computeVelocity isoThermalDynamics.hh:1389-1396: u = j * invRho0
            T dummyRhoBar;
            get_rhoBar_j(f, dummyRhoBar, u);
            u[0] *= (T)1; u[1] *= (T)1; u[2] *= (T)1;
            return;
        }
        get_j(f, u);
        const T invRho = invRhoOf(get_rhoBar(f));
        u[0] *= invRho; u[1] *= invRho; u[2] *= invRho;
    }
    PLB_HD static bool velocity_reads_pops(int dynType, bool hasBc, int bcKind) {
        if (hasBc) return is_rho_plane(bcKind);
        return !is_inert(dynType);
    }

    /// Dynamics::computeRhoBar: bulk core/dynamics.hh:351-354; BB :861-863; velocity planes
    /// boundaryDynamics.hh:639-668; density planes :254-257; StoreVelocityDynamics :342-346.
    PLB_HD static T cell_rhoBar(const Dyn<T>& d, bool hasBc, const Bc<T>& b, const T* f) {
        if (hasBc) {
            if (is_vel_plane(b.kind)) return veldir_rhoBar(b, f, vel_is_j(d.type));
            if (is_rho_plane(b.kind)) return b.rhoBar;
            if (is_inner(b.kind)) return inner_rhoBar(b, f, vel_is_j(d.type));
            return rhoBarOf(cell_density(d, hasBc, b, f));
        }
        if (is_inert(d.type)) return rhoBarOf(d.rho);
        return get_rhoBar(f);
    }

    /// Dynamics::computeRhoBarJ of whatever sits on the cell -- what BoxRhoBarJfunctional3D stores
    /// (dataProcessors/dataAnalysisFunctional3D.hh:528-548): bulk get_rhoBar_j core/dynamics.hh:357-361; BounceBack rhoBar(rho), 0
    /// :880-885; NoDynamics 1 - SkordosFactor, 0 :1061-1066; velocity planes boundaryDynamics.hh:682-701; density planes :816-822;
    /// StoreVelocityDynamics (edges, corners) rho = computeDensity, j = rho * velocity :348-359
    PLB_HD static void cell_rhoBar_j(const Dyn<T>& d, bool hasBc, const Bc<T>& b, const T* f, T& rhoBar, T j[3]) {
        if (hasBc) {
            if (b.kind == BC_STORE_VEL) {
                const T rho = cell_density(d, hasBc, b, f);
                rhoBar = rhoBarOf(rho);
#pragma unroll
                for (int k = 0; k < 3; ++k) j[k] = rho * (b.u[k] - 0.5 * (T)0);
            } else if (is_inner(b.kind)) {
                inner_rhoBar_j(b, f, rhoBar, j, vel_is_j(d.type));
            } else {
                plane_rhoBar_j(b, f, rhoBar, j, vel_is_j(d.type));
            }
            return;
        }
        if (d.type == DYN_BOUNCEBACK) { rhoBar = rhoBarOf(d.rho); j[0] = j[1] = j[2] = (T)0; return; }
        if (d.type == DYN_NODYNAMICS) { rhoBar = (T)1 - (T)1; j[0] = j[1] = j[2] = (T)0; return; }
        get_rhoBar_j(f, rhoBar, j);
    }
    PLB_HD static bool rhoBar_j_reads_pops(int dynType, bool hasBc, int bcKind) {
        if (hasBc) return bcKind == BC_STORE_VEL ? !is_inert(dynType) : true;
        return !is_inert(dynType);
    }

    /// iniCellAtEquilibrium core/cell.hh:124-132
    PLB_HD static void ini_equilibrium(int type, T* f, T density, const T velocity[3]) {
        T j[3];
        j[0] = velocity[0] * density; j[1] = velocity[1] * density; j[2] = velocity[2] * density;
        const T jSqr = j[0] * j[0] + j[1] * j[1] + j[2] * j[2];
        const T rhoBar = rhoBarOf(density);
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = computeEquilibrium(type, i, rhoBar, j, jSqr);
    }
};

}  // namespace plb200
