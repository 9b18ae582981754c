/* oracle/plb_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU oracle: a plain-C restatement of the reference's (Palabos,
 * /root/reference/src) collide-and-stream path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (libplb_b200.so) never does.  Parity pin: tests/test_oracle_golden.py compares it
 * bit-for-bit with dumps of the real reference (oracle/_ref/ref_driver) committed under
 * tests/golden/, and with the known-answer strings of the reference's own regression
 * test jlabos/src/testing/test_cylinder.py:85-97.
 *
 * The entry points mirror include/plb_b200.h one to one (prefix plbo_ instead of
 * plb_b200_), so the same scenario can be replayed on both.  Return convention as
 * there: 1 = success, 0 = failure (coProcessors/coProcessor3D.h:77).
 */
#ifndef PLB_ORACLE_H
#define PLB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* numeric ids shared with include/plb_b200.h */
enum { PLBO_DYN_BGK = 0, PLBO_DYN_RLB = 1, PLBO_DYN_TRT = 2, PLBO_DYN_SMAGO = 3,
       PLBO_DYN_BOUNCEBACK = 4, PLBO_DYN_NODYNAMICS = 5, PLBO_DYN_INCBGK = 6 };
enum { PLBO_BC_LOCAL = 0, PLBO_BC_ZOUHE = 1, PLBO_BC_DYNLOCAL = 2, PLBO_BC_EQUILIBRIUM = 3, PLBO_BC_INTERP = 4 };

typedef struct plbo_lattice plbo_lattice;

plbo_lattice* plbo_create(int q, int precision_bytes, int nx, int ny, int nz,
                          int dynamics, double p1, double p2);
void plbo_destroy(plbo_lattice* L);
const char* plbo_last_error(void);

int plbo_set_periodicity(plbo_lattice* L, int px, int py, int pz);
/* Cells outside `box` neither gather statistics nor run boundary processors: they play the
 * role of the reference's envelope cells (multiBlockLattice3D.hh:403-421). */
int plbo_set_statistics_box(plbo_lattice* L, const int box[6]);
int plbo_define_dynamics_box(plbo_lattice* L, const int box[6], int dynamics, double p1, double p2);
int plbo_define_dynamics_mask(plbo_lattice* L, const int32_t* mask, int flag, int dynamics, double p1, double p2);
int plbo_add_velocity_boundary(plbo_lattice* L, int model, int direction, int orientation, const int box[6]);
int plbo_add_pressure_boundary(plbo_lattice* L, int model, int direction, int orientation, const int box[6]);
int plbo_add_external_velocity_edge(plbo_lattice* L, int model, int plane, int normal1, int normal2, const int box[6]);
int plbo_add_external_velocity_corner(plbo_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z);
/* addInternalVelocityEdge{0,1,2}{N,P}{N,P} / addInternalVelocityCorner{N,P}{N,P}{N,P}: the concave edges and corners */
int plbo_add_internal_velocity_edge(plbo_lattice* L, int model, int plane, int normal1, int normal2, const int box[6]);
int plbo_add_internal_velocity_corner(plbo_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z);
int plbo_set_velocity_condition_on_block_boundaries(plbo_lattice* L, int model, const int block[6], const int app[6]);
int plbo_set_boundary_velocity(plbo_lattice* L, const int box[6], const double u[3]);
int plbo_set_boundary_density(plbo_lattice* L, const int box[6], double rho);
int plbo_initialize_at_equilibrium(plbo_lattice* L, const int box[6], double rho, const double u[3]);
int plbo_initialize_at_equilibrium_field(plbo_lattice* L, const int box[6], const double* rho_u);
int plbo_upload(plbo_lattice* L, const int box[6], const void* data, size_t bytes);
int plbo_download(plbo_lattice* L, const int box[6], void* data, size_t bytes);
int plbo_initialize(plbo_lattice* L);
int plbo_collide_and_stream(plbo_lattice* L, int n_steps);
int plbo_get_stored_statistics(plbo_lattice* L, double out3[3], uint64_t* n_cells);
int plbo_compute_average_energy(plbo_lattice* L, const int box[6], double* out);
int plbo_compute_average_density(plbo_lattice* L, const int box[6], double* out);
int plbo_compute_density_velocity(plbo_lattice* L, const int box[6], void* rho_u, size_t bytes);

/* External-moment sweep: ExternalRhoJcollideAndStream3D at level 0 + BoxRhoBarJfunctional3D at
 * level 3 (basicDynamics/dynamicsProcessor3D.hh:41-195, dataProcessors/dataAnalysisFunctional3D.hh:528-548).  The lattice
 * must have been created one cell larger on every non-periodic side, with plbo_set_statistics_box = the bulk: the
 * extra layer is the reference's envelope, which this sweep collides and streams (tests/scenario.py does the shift). */
int plbo_enable_external_rhoj(plbo_lattice* L);
int plbo_execute_internal_processors(plbo_lattice* L, int n_steps);
int plbo_apply_rhobar_j(plbo_lattice* L, const int box[6]);
int plbo_upload_rhobar_j(plbo_lattice* L, const int box[6], const void* data, size_t bytes);
int plbo_download_rhobar_j(plbo_lattice* L, const int box[6], void* data, size_t bytes);

/* per-cell functors, exported so that unit tests can check the device functors one by one */
int plbo_collide_cell(int q, int precision_bytes, int dynamics, double p1, double p2,
                      void* f, double* rhoBar_uSqr);

#ifdef __cplusplus
}
#endif
#endif
