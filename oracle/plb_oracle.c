/* oracle/plb_oracle.c -- TEST INFRASTRUCTURE ONLY (see plb_oracle.h for the contract).
 *
 * Plain-C CPU restatement of the reference's hot path (Palabos MultiBlockLattice3D::
 * collideAndStream with its boundary processors and statistics), for D3Q19 / D3Q27 in
 * float and double.  The arithmetic lives in plb_oracle_impl.h (included once per
 * precision); this file holds the descriptor tables, the set-up calls that mirror the
 * reference's wrappers, and the precision dispatch.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "plb_oracle.h"

#define MAX_DYN 128
enum { BC_VEL_REG = 0, BC_RHO_REG = 1, BC_VEL_ZH = 2, BC_RHO_ZH = 3, BC_STORE_VEL = 4, BC_INNER_EDGE = 5, BC_INNER_CORNER = 6 };
enum { PROC_PLANE = 0, PROC_EDGE = 1, PROC_CORNER = 2, PROC_PLANE_FD = 3 };
typedef struct { int cls; int box[6]; int a, b, c; } Proc;

/* latticeBoltzmann/nearestNeighborLattices3D.hh:165-177 (D3Q19 c), :225-239 (D3Q27 c) */
static const int c19[19][3] = {
    { 0, 0, 0},
    {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1},
    {-1,-1, 0}, {-1, 1, 0}, {-1, 0,-1},
    {-1, 0, 1}, { 0,-1,-1}, { 0,-1, 1},
    { 1, 0, 0}, { 0, 1, 0}, { 0, 0, 1},
    { 1, 1, 0}, { 1,-1, 0}, { 1, 0, 1},
    { 1, 0,-1}, { 0, 1, 1}, { 0, 1,-1} };
static const int c19n[19] = { 0, 1,1,1, 2,2,2,2,2,2, 1,1,1, 2,2,2,2,2,2 };
static const int c27[27][3] = {
    { 0, 0, 0},
    {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1},
    {-1,-1, 0}, {-1, 1, 0}, {-1, 0,-1},
    {-1, 0, 1}, { 0,-1,-1}, { 0,-1, 1},
    {-1,-1,-1}, {-1,-1, 1}, {-1, 1,-1}, {-1, 1, 1},
    { 1, 0, 0}, { 0, 1, 0}, { 0, 0, 1},
    { 1, 1, 0}, { 1,-1, 0}, { 1, 0, 1},
    { 1, 0,-1}, { 0, 1, 1}, { 0, 1,-1},
    { 1, 1, 1}, { 1, 1,-1}, { 1,-1, 1}, { 1,-1,-1} };
static const int c27n[27] = { 0, 1,1,1, 2,2,2,2,2,2, 3,3,3,3, 1,1,1, 2,2,2,2,2,2, 3,3,3,3 };

#define REAL double
#define SFX(x) x##_f64
#define SQRT sqrt
#include "plb_oracle_impl.h"
#undef REAL
#undef SFX
#undef SQRT

#define REAL float
#define SFX(x) x##_f32
#define SQRT sqrtf
#include "plb_oracle_impl.h"
#undef REAL
#undef SFX
#undef SQRT

struct plbo_lattice { int prec; Lat_f64* d; Lat_f32* s; };

static char g_err[256] = "";
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return 0; }
const char* plbo_last_error(void) { return g_err; }

#define DISPATCH(L, expr64, expr32) ((L)->prec == 8 ? (expr64) : (expr32))
#define NX(L) DISPATCH(L, (L)->d->nx, (L)->s->nx)
#define NY(L) DISPATCH(L, (L)->d->ny, (L)->s->ny)
#define NZ(L) DISPATCH(L, (L)->d->nz, (L)->s->nz)
#define QQ(L) DISPATCH(L, (L)->d->q,  (L)->s->q)

static int box_ok(plbo_lattice* L, const int b[6]) {
    return b[0] >= 0 && b[1] < NX(L) && b[2] >= 0 && b[3] < NY(L) && b[4] >= 0 && b[5] < NZ(L)
        && b[0] <= b[1] && b[2] <= b[3] && b[4] <= b[5];
}
static size_t box_cells(const int b[6]) {
    return (size_t)(b[1]-b[0]+1) * (b[3]-b[2]+1) * (b[5]-b[4]+1);
}

plbo_lattice* plbo_create(int q, int prec, int nx, int ny, int nz, int dynamics, double p1, double p2) {
    plbo_lattice* L;
    if ((q != 19 && q != 27) || (prec != 4 && prec != 8) || nx < 1 || ny < 1 || nz < 1) { fail("bad lattice descriptor"); return 0; }
    L = (plbo_lattice*)calloc(1, sizeof *L);
    L->prec = prec;
    if (prec == 8) { L->d = create_f64(q, nx, ny, nz); find_dyn_f64(L->d, dynamics, p1, p2); }
    else           { L->s = create_f32(q, nx, ny, nz); find_dyn_f32(L->s, dynamics, p1, p2); }
    return L;
}

void plbo_destroy(plbo_lattice* L) {
    if (!L) return;
    if (L->d) destroy_f64(L->d);
    if (L->s) destroy_f32(L->s);
    free(L);
}

int plbo_set_periodicity(plbo_lattice* L, int px, int py, int pz) {
    int* per = DISPATCH(L, L->d->per, L->s->per);
    per[0] = px != 0; per[1] = py != 0; per[2] = pz != 0;
    return 1;
}

int plbo_set_statistics_box(plbo_lattice* L, const int box[6]) {
    if (!box_ok(L, box)) return fail("box outside the lattice");
    memcpy(DISPATCH(L, L->d->statBox, L->s->statBox), box, 6 * sizeof(int));
    return 1;
}

/* defineDynamics(lattice, box, dynamics) dataProcessors/dataInitializerWrapper3D.h:163-165:
 * the cell's whole dynamics object is replaced, so a boundary composite on it disappears. */
static int define_cell(plbo_lattice* L, size_t cell, int idx) {
    if (idx < 0) return fail("too many distinct dynamics");
    if (L->prec == 8) { L->d->dynIdx[cell] = (uint8_t)idx; L->d->bcIdx[cell] = -1; }
    else              { L->s->dynIdx[cell] = (uint8_t)idx; L->s->bcIdx[cell] = -1; }
    return 1;
}

int plbo_define_dynamics_box(plbo_lattice* L, const int box[6], int dynamics, double p1, double p2) {
    int x, y, z, idx;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    idx = DISPATCH(L, find_dyn_f64(L->d, dynamics, p1, p2), find_dyn_f32(L->s, dynamics, p1, p2));
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z)
        if (!define_cell(L, ((size_t)x * NY(L) + y) * NZ(L) + z, idx)) return 0;
    return 1;
}

/* defineDynamics(lattice, intMask, dynamics, flag) dataInitializerWrapper3D.h:194-200 */
int plbo_define_dynamics_mask(plbo_lattice* L, const int32_t* mask, int flag, int dynamics, double p1, double p2) {
    size_t n = (size_t)NX(L) * NY(L) * NZ(L), k;
    int idx = DISPATCH(L, find_dyn_f64(L->d, dynamics, p1, p2), find_dyn_f32(L->s, dynamics, p1, p2));
    for (k = 0; k < n; ++k) if (mask[k] == flag) if (!define_cell(L, k, idx)) return 0;
    return 1;
}

static int add_bc(plbo_lattice* L, int kind, int dir, int ori, const int box[6]) {
    if (!DISPATCH(L, add_bc_f64(L->d, kind, dir, ori, box), add_bc_f32(L->s, kind, dir, ori, box)))
        return fail("cell already carries a boundary dynamics");
    return 1;
}
static void add_proc(plbo_lattice* L, int cls, const int box[6], int a, int b, int c) {
    if (L->prec == 8) add_proc_f64(L->d, cls, box, a, b, c); else add_proc_f32(L->s, cls, box, a, b, c);
}

/* Who completes the populations of a plane boundary cell.  Wrapped managers (createLocalBoundaryCondition3D,
 * createZouHeBoundaryCondition3D): a WrappedLocalBoundaryFunctional3D processor at level 0, after streaming.  Dynamics-based
 * managers (createDynamicsBasedLocalBoundaryCondition3D, createEquilibriumBoundaryCondition3D, boundaryCondition3D.hh:1170-1181:
 * in this fork both hand out the regularized dynamics and NO functional, :712-747, :940-974): the dynamics itself at the
 * beginning of collide() (automaticPrepareCollision, core/dynamics.hh:470-475). */
static void plane_completion(plbo_lattice* L, int model, int direction, int orientation, const int box[6]) {
    /* createInterpBoundaryCondition3D (boundaryCondition3D.hh:1049-1085): plain Dirichlet dynamics + PlaneFdBoundaryFunctional3D */
    if (model == PLBO_BC_INTERP) { add_proc(L, PROC_PLANE_FD, box, direction, orientation, 0); return; }
    if (model >= PLBO_BC_DYNLOCAL) { if (L->prec == 8) mark_auto_f64(L->d, box); else mark_auto_f32(L->s, box); }
    else add_proc(L, PROC_PLANE, box, 0, 0, 0);
}

/* BoundaryConditionInstantiator3D::addVelocityBoundary boundaryInstantiator3D.h:1016-1070 with
 * WrappedRegularizedBoundaryManager3D (boundaryCondition3D.hh:825-840) or
 * WrappedZouHeBoundaryManager3D (zouHeBoundary3D.hh): composite dynamics on the cells, then the
 * WrappedLocalBoundaryFunctional3D processor at level 0.  bcType = dirichlet only. */
int plbo_add_velocity_boundary(plbo_lattice* L, int model, int direction, int orientation, const int box[6]) {
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (direction < 0 || direction > 2 || (orientation != 1 && orientation != -1)) return fail("bad direction/orientation");
    if (model < 0 || model > PLBO_BC_INTERP) return fail("unknown boundary condition model");
    if (!add_bc(L, model == PLBO_BC_ZOUHE ? BC_VEL_ZH : BC_VEL_REG, direction, orientation, box)) return 0;
    plane_completion(L, model, direction, orientation, box);
    return 1;
}
int plbo_add_pressure_boundary(plbo_lattice* L, int model, int direction, int orientation, const int box[6]) {
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (direction < 0 || direction > 2 || (orientation != 1 && orientation != -1)) return fail("bad direction/orientation");
    if (model < 0 || model > PLBO_BC_INTERP) return fail("unknown boundary condition model");
    if (!add_bc(L, model == PLBO_BC_ZOUHE ? BC_RHO_ZH : BC_RHO_REG, direction, orientation, box)) return 0;
    plane_completion(L, model, direction, orientation, box);
    return 1;
}
/* addExternalVelocityEdge / Corner: StoreVelocityDynamics + Outer{Edge,Corner}Functional3D
 * (boundaryCondition3D.hh:861-877, :899-915; same for the Zou-He manager) */
int plbo_add_external_velocity_edge(plbo_lattice* L, int model, int plane, int n1, int n2, const int box[6]) {
    (void)model;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (!add_bc(L, BC_STORE_VEL, 0, 0, box)) return 0;
    add_proc(L, PROC_EDGE, box, plane, n1, n2);
    return 1;
}
int plbo_add_external_velocity_corner(plbo_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z) {
    int box[6] = { x, x, y, y, z, z };
    (void)model;
    if (!box_ok(L, box)) return fail("corner outside the lattice");
    if (!add_bc(L, BC_STORE_VEL, 0, 0, box)) return 0;
    add_proc(L, PROC_CORNER, box, xn, yn, zn);
    return 1;
}

/* addInternalVelocityEdge / Corner (boundaryInstantiator3D.h; managers boundaryCondition3D.hh:770-820 and its wrapped / Zou-He /
 * equilibrium / interpolation counterparts): RegularizedVelocityInnerEdge/CornerDynamics3D on the cells; the wrapped managers
 * (local, Zou-He) add the WrappedLocalBoundaryFunctional3D processor, the others leave automaticPrepareCollision on. */
int plbo_add_internal_velocity_edge(plbo_lattice* L, int model, int plane, int n1, int n2, const int box[6]) {
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (plane < 0 || plane > 2 || (n1 != 1 && n1 != -1) || (n2 != 1 && n2 != -1)) return fail("bad plane / normals");
    if (model < 0 || model > PLBO_BC_INTERP) return fail("unknown boundary condition model");
    if (!add_bc(L, BC_INNER_EDGE, plane, (n1 > 0 ? 2 : 0) + (n2 > 0 ? 1 : 0), box)) return 0;
    if (model >= PLBO_BC_DYNLOCAL) { if (L->prec == 8) mark_auto_f64(L->d, box); else mark_auto_f32(L->s, box); }
    else add_proc(L, PROC_PLANE, box, 0, 0, 0);
    return 1;
}
int plbo_add_internal_velocity_corner(plbo_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z) {
    int box[6] = { x, x, y, y, z, z };
    if (!box_ok(L, box)) return fail("corner outside the lattice");
    if (model < 0 || model > PLBO_BC_INTERP) return fail("unknown boundary condition model");
    if (!add_bc(L, BC_INNER_CORNER, 0, (xn > 0 ? 4 : 0) + (yn > 0 ? 2 : 0) + (zn > 0 ? 1 : 0), box)) return 0;
    if (model >= PLBO_BC_DYNLOCAL) { if (L->prec == 8) mark_auto_f64(L->d, box); else mark_auto_f32(L->s, box); }
    else add_proc(L, PROC_PLANE, box, 0, 0, 0);
    return 1;
}

static int intersect(const int a[6], const int b[6], int out[6]) {
    int k;
    for (k = 0; k < 3; ++k) {
        out[2*k]   = a[2*k]   > b[2*k]   ? a[2*k]   : b[2*k];
        out[2*k+1] = a[2*k+1] < b[2*k+1] ? a[2*k+1] : b[2*k+1];
        if (out[2*k] > out[2*k+1]) return 0;
    }
    return 1;
}

/* OnLatticeBoundaryCondition3D::setVelocityConditionOnBlockBoundaries
 * boundaryCondition/boundaryCondition3D.hh:64-162 with BlockSurface3D(block, 1)
 * (core/blockSurface3D.cpp:38-200): 6 faces, 12 edges, 8 corners, in that order. */
int plbo_set_velocity_condition_on_block_boundaries(plbo_lattice* L, int model, const int d[6], const int app[6]) {
    const int x0 = d[0], x1 = d[1], y0 = d[2], y1 = d[3], z0 = d[4], z1 = d[5];
    int in[6], k;
    /* faces: dir, ori, box */
    const int faces[6][8] = {
        {0,-1, x0,x0,     y0+1,y1-1, z0+1,z1-1}, {0,+1, x1,x1,     y0+1,y1-1, z0+1,z1-1},
        {1,-1, x0+1,x1-1, y0,y0,     z0+1,z1-1}, {1,+1, x0+1,x1-1, y1,y1,     z0+1,z1-1},
        {2,-1, x0+1,x1-1, y0+1,y1-1, z0,z0},     {2,+1, x0+1,x1-1, y0+1,y1-1, z1,z1} };
    /* edges: plane, normal1 (along (plane+1)%3), normal2 (along (plane+2)%3), box */
    const int edges[12][9] = {
        {0,-1,-1, x0+1,x1-1, y0,y0, z0,z0}, {0,-1,+1, x0+1,x1-1, y0,y0, z1,z1},
        {0,+1,-1, x0+1,x1-1, y1,y1, z0,z0}, {0,+1,+1, x0+1,x1-1, y1,y1, z1,z1},
        {1,-1,-1, x0,x0, y0+1,y1-1, z0,z0}, {1,-1,+1, x1,x1, y0+1,y1-1, z0,z0},
        {1,+1,-1, x0,x0, y0+1,y1-1, z1,z1}, {1,+1,+1, x1,x1, y0+1,y1-1, z1,z1},
        {2,-1,-1, x0,x0, y0,y0, z0+1,z1-1}, {2,-1,+1, x0,x0, y1,y1, z0+1,z1-1},
        {2,+1,-1, x1,x1, y0,y0, z0+1,z1-1}, {2,+1,+1, x1,x1, y1,y1, z0+1,z1-1} };
    for (k = 0; k < 6; ++k)
        if (intersect(&faces[k][2], app, in))
            if (!plbo_add_velocity_boundary(L, model, faces[k][0], faces[k][1], in)) return 0;
    for (k = 0; k < 12; ++k)
        if (intersect(&edges[k][3], app, in))
            if (!plbo_add_external_velocity_edge(L, model, edges[k][0], edges[k][1], edges[k][2], in)) return 0;
    for (k = 0; k < 8; ++k) {
        const int xn = (k & 4) ? 1 : -1, yn = (k & 2) ? 1 : -1, zn = (k & 1) ? 1 : -1;
        const int cb[6] = { xn < 0 ? x0 : x1, xn < 0 ? x0 : x1, yn < 0 ? y0 : y1, yn < 0 ? y0 : y1,
                            zn < 0 ? z0 : z1, zn < 0 ? z0 : z1 };
        if (intersect(cb, app, in))
            if (!plbo_add_external_velocity_corner(L, model, xn, yn, zn, in[0], in[2], in[4])) return 0;
    }
    return 1;
}

/* setBoundaryVelocity dataInitializerFunctional3D.hh:830-845 -> Cell::defineVelocity: only the
 * StoreVelocity-type dynamics keep it (boundaryDynamics.hh:333-339); no-op elsewhere. */
int plbo_set_boundary_velocity(plbo_lattice* L, const int box[6], const double u[3]) {
    int x, y, z;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z) {
        size_t cell = ((size_t)x * NY(L) + y) * NZ(L) + z;
        if (L->prec == 8) {
            int bi = L->d->bcIdx[cell];
            if (bi >= 0 && L->d->bcs[bi].kind != BC_RHO_REG && L->d->bcs[bi].kind != BC_RHO_ZH) {
                L->d->bcs[bi].u[0] = u[0]; L->d->bcs[bi].u[1] = u[1]; L->d->bcs[bi].u[2] = u[2];
            }
        } else {
            int bi = L->s->bcIdx[cell];
            if (bi >= 0 && L->s->bcs[bi].kind != BC_RHO_REG && L->s->bcs[bi].kind != BC_RHO_ZH) {
                L->s->bcs[bi].u[0] = (float)u[0]; L->s->bcs[bi].u[1] = (float)u[1]; L->s->bcs[bi].u[2] = (float)u[2];
            }
        }
    }
    return 1;
}

/* setBoundaryDensity -> Cell::defineDensity: StoreDensityDynamics::defineDensity
 * boundaryDynamics.hh:239-243 (rhoBar = rho - 1) */
int plbo_set_boundary_density(plbo_lattice* L, const int box[6], double rho) {
    int x, y, z;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z) {
        size_t cell = ((size_t)x * NY(L) + y) * NZ(L) + z;
        if (L->prec == 8) {
            int bi = L->d->bcIdx[cell];
            if (bi >= 0 && (L->d->bcs[bi].kind == BC_RHO_REG || L->d->bcs[bi].kind == BC_RHO_ZH))
                L->d->bcs[bi].rhoBar = rho - 1.0;
        } else {
            int bi = L->s->bcIdx[cell];
            if (bi >= 0 && (L->s->bcs[bi].kind == BC_RHO_REG || L->s->bcs[bi].kind == BC_RHO_ZH))
                L->s->bcs[bi].rhoBar = (float)rho - 1.0f;
        }
    }
    return 1;
}

/* initializeAtEquilibrium(lattice, box, rho, u) -> iniCellAtEquilibrium core/cell.hh:124-132 */
int plbo_initialize_at_equilibrium(plbo_lattice* L, const int box[6], double rho, const double u[3]) {
    int x, y, z;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z) {
        size_t cell = ((size_t)x * NY(L) + y) * NZ(L) + z;
        if (L->prec == 8) { double v[3] = { u[0], u[1], u[2] }; ini_equilibrium_f64(L->d, cell, rho, v); }
        else { float v[3] = { (float)u[0], (float)u[1], (float)u[2] }; ini_equilibrium_f32(L->s, cell, (float)rho, v); }
    }
    return 1;
}

/* initializeAtEquilibrium(lattice, box, RhoUFunction) (in-place overload,
 * dataInitializerWrapper3D.h): rho_u = [cells of box in x,y,z order][rho,ux,uy,uz] */
int plbo_initialize_at_equilibrium_field(plbo_lattice* L, const int box[6], const double* ru) {
    int x, y, z;
    size_t k = 0;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z, ++k) {
        size_t cell = ((size_t)x * NY(L) + y) * NZ(L) + z;
        if (L->prec == 8) { double v[3] = { ru[4*k+1], ru[4*k+2], ru[4*k+3] }; ini_equilibrium_f64(L->d, cell, ru[4*k], v); }
        else { float v[3] = { (float)ru[4*k+1], (float)ru[4*k+2], (float)ru[4*k+3] }; ini_equilibrium_f32(L->s, cell, (float)ru[4*k], v); }
    }
    return 1;
}

/* wire format of CoProcessor3D::send/receive coProcessors/coProcessor3D.h:79-92: raw T,
 * [x][y][z][q], values as stored (f - t_i), i.e. Cell::serialize core/cell.hh:103-111 */
static int transfer(plbo_lattice* L, const int box[6], void* data, size_t bytes, int up) {
    const int q = QQ(L);
    int x, y, z;
    size_t k = 0, esz = (size_t)L->prec;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (bytes != box_cells(box) * q * esz) return fail("buffer size does not match the box");
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z, ++k) {
        size_t cell = ((size_t)x * NY(L) + y) * NZ(L) + z;
        char* lat = L->prec == 8 ? (char*)(L->d->f + cell * q) : (char*)(L->s->f + cell * q);
        if (up) memcpy(lat, (char*)data + k * q * esz, q * esz);
        else    memcpy((char*)data + k * q * esz, lat, q * esz);
    }
    return 1;
}
int plbo_upload(plbo_lattice* L, const int box[6], const void* data, size_t bytes) { return transfer(L, box, (void*)data, bytes, 1); }
int plbo_download(plbo_lattice* L, const int box[6], void* data, size_t bytes) { return transfer(L, box, data, bytes, 0); }

/* lattice.initialize(): MultiBlock3D::initialize -> executeInternalProcessors (multiBlock3D.cpp) */
int plbo_initialize(plbo_lattice* L) {
    if (L->prec == 8) execute_processors_f64(L->d); else execute_processors_f32(L->s);
    return 1;
}

int plbo_collide_and_stream(plbo_lattice* L, int n) {
    int k;
    for (k = 0; k < n; ++k) { if (L->prec == 8) step_f64(L->d); else step_f32(L->s); }
    return 1;
}

/* ---- the external-moment sweep, see ext_step in plb_oracle_impl.h ---- */

/* generateMultiScalarField / generateMultiTensorField<3> + the two integrateProcessingFunctional calls, src/variables.hh:475-486 */
int plbo_enable_external_rhoj(plbo_lattice* L) {
    size_t n = (size_t)NX(L) * NY(L) * NZ(L);
    if (L->prec == 8) { if (!L->d->ext) L->d->ext = (double*)calloc(4 * n, sizeof(double)); }
    else              { if (!L->s->ext) L->s->ext = (float*)calloc(4 * n, sizeof(float)); }
    return 1;
}
static int has_ext(plbo_lattice* L) { return DISPATCH(L, L->d->ext != 0, L->s->ext != 0); }

/* n x { lattice.executeInternalProcessors(); lattice.incrementTime(); } viscosityTest/src/test.cpp:64-65 */
int plbo_execute_internal_processors(plbo_lattice* L, int n) {
    int k;
    if (!has_ext(L)) return fail("external rhoBar / j fields are not enabled");
    for (k = 0; k < n; ++k) { if (L->prec == 8) ext_step_f64(L->d); else ext_step_f32(L->s); }
    return 1;
}

/* applyProcessingFunctional(new BoxRhoBarJfunctional3D, box, rhoBarJarg) src/variables.hh:650 */
int plbo_apply_rhobar_j(plbo_lattice* L, const int box[6]) {
    if (!has_ext(L)) return fail("external rhoBar / j fields are not enabled");
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (L->prec == 8) box_rhoBar_j_f64(L->d, box); else box_rhoBar_j_f32(L->s, box);
    return 1;
}

/* the fields as T [cells of box][rhoBar, jx, jy, jz] */
static int transfer_ext(plbo_lattice* L, const int box[6], void* data, size_t bytes, int up) {
    int x, y, z;
    size_t k = 0, esz = (size_t)L->prec;
    if (!has_ext(L)) return fail("external rhoBar / j fields are not enabled");
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (bytes != box_cells(box) * 4 * esz) return fail("buffer size does not match the box");
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z, ++k) {
        size_t cell = ((size_t)x * NY(L) + y) * NZ(L) + z;
        char* fld = L->prec == 8 ? (char*)(L->d->ext + cell * 4) : (char*)(L->s->ext + cell * 4);
        if (up) memcpy(fld, (char*)data + k * 4 * esz, 4 * esz);
        else    memcpy((char*)data + k * 4 * esz, fld, 4 * esz);
    }
    return 1;
}
int plbo_upload_rhobar_j(plbo_lattice* L, const int box[6], const void* data, size_t bytes) { return transfer_ext(L, box, (void*)data, bytes, 1); }
int plbo_download_rhobar_j(plbo_lattice* L, const int box[6], void* data, size_t bytes) { return transfer_ext(L, box, data, bytes, 0); }

/* out3 = { average rhoBar, average uSqr, max uSqr } (core/latticeStatistics.h:32-34) */
int plbo_get_stored_statistics(plbo_lattice* L, double out3[3], uint64_t* n) {
    if (L->prec == 8) { out3[0] = L->d->avRhoBar; out3[1] = L->d->avUSqr; out3[2] = L->d->maxUSqr; if (n) *n = L->d->numCells; }
    else              { out3[0] = L->s->avRhoBar; out3[1] = L->s->avUSqr; out3[2] = L->s->maxUSqr; if (n) *n = L->s->numCells; }
    return 1;
}

int plbo_compute_average_energy(plbo_lattice* L, const int box[6], double* out) {
    if (!box_ok(L, box)) return fail("box outside the lattice");
    *out = DISPATCH(L, average_energy_f64(L->d, box), average_energy_f32(L->s, box));
    return 1;
}
int plbo_compute_average_density(plbo_lattice* L, const int box[6], double* out) {
    if (!box_ok(L, box)) return fail("box outside the lattice");
    *out = DISPATCH(L, average_density_f64(L->d, box), average_density_f32(L->s, box));
    return 1;
}

/* rho_u = [cells][rho,ux,uy,uz] in T through Cell::computeDensity / computeVelocity */
int plbo_compute_density_velocity(plbo_lattice* L, const int box[6], void* out, size_t bytes) {
    int x, y, z;
    size_t k = 0;
    if (!box_ok(L, box)) return fail("box outside the lattice");
    if (bytes != box_cells(box) * 4 * (size_t)L->prec) return fail("buffer size does not match the box");
    for (x = box[0]; x <= box[1]; ++x) for (y = box[2]; y <= box[3]; ++y) for (z = box[4]; z <= box[5]; ++z, ++k) {
        size_t cell = ((size_t)x * NY(L) + y) * NZ(L) + z;
        if (L->prec == 8) { double* o = (double*)out + 4*k; o[0] = cell_density_f64(L->d, cell); cell_velocity_f64(L->d, cell, o + 1); }
        else              { float*  o = (float*)out + 4*k;  o[0] = cell_density_f32(L->s, cell); cell_velocity_f32(L->s, cell, o + 1); }
    }
    return 1;
}

/* one Dynamics::collide on a bare population array, for the functor unit tests */
int plbo_collide_cell(int q, int prec, int dynamics, double p1, double p2, void* f, double* out2) {
    int took;
    if (prec == 8) {
        Lat_f64* L = create_f64(q, 1, 1, 1);
        double r = 0, u = 0;
        find_dyn_f64(L, dynamics, p1, p2);
        took = collide_cell_f64(L, &L->dyns[0], (double*)f, &r, &u);
        if (out2) { out2[0] = r; out2[1] = u; }
        destroy_f64(L);
    } else {
        Lat_f32* L = create_f32(q, 1, 1, 1);
        float r = 0, u = 0;
        find_dyn_f32(L, dynamics, p1, p2);
        took = collide_cell_f32(L, &L->dyns[0], (float*)f, &r, &u);
        if (out2) { out2[0] = r; out2[1] = u; }
        destroy_f32(L);
    }
    return took;
}
