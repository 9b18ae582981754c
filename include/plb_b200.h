/* include/plb_b200.h -- C-ABI of libplb_b200.so, the B200 engine for ONE hot path of Palabos:
 * MultiBlockLattice3D::collideAndStream() with its boundary processors, envelope (halo) exchange
 * and BlockStatistics.  Citations are relative to the reference tree (/root/reference/src).
 *
 * Conventions (as the reference's own accelerator plug-in, coProcessors/coProcessor3D.h:59-116):
 *   - every call returns 1 on success, 0 on failure (coProcessor3D.h:77,93,103,115); nothing throws
 *     across this boundary; plb_b200_last_error() tells why (thread-local string);
 *   - plain pointers and sizes only; a lattice is an opaque handle; all calls on one lattice come
 *     from one host thread; kernels are stream-ordered, calls block only where they return data;
 *   - boxes are int[6] = {x0,x1,y0,y1,z0,z1}, inclusive, GLOBAL lattice coordinates (plb::Box3D);
 *   - populations cross the boundary in the wire format of CoProcessor3D::send/receive
 *     (coProcessor3D.h:79-92): raw T, [x][y][z][q], q fastest, D3Q19Constants / D3Q27Constants
 *     order (latticeBoltzmann/nearestNeighborLattices3D.hh:159-196, :219-262), values as stored by
 *     Cell::serialize (core/cell.hh:103-111), i.e. f_i - t_i;
 *   - there is no CPU fallback: without a CUDA device plb_b200_create fails.
 *
 * A header-only C++ mirror of the reference's classes on top of this ABI is include/palabos_b200/.
 * How a Palabos maintainer binds it: INTEGRATION.md.
 */
#ifndef PLB_B200_H
#define PLB_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLB_B200_ABI_VERSION 1

/* Dynamics that can sit on a cell (Dynamics::collide, core/dynamics.h:102-103):
 *   BGK        BGKdynamics              basicDynamics/isoThermalDynamics.hh:319-331     p1 = omega
 *   RLB        RegularizedBGKdynamics   basicDynamics/isoThermalDynamics.hh:1588-1602   p1 = omega
 *   TRT        TRTdynamics              complexDynamics/trtDynamics.hh:64-99            p1 = omega (s- = 1.1)
 *   SMAGO      SmagorinskyBGKdynamics   complexDynamics/smagorinskyDynamics.hh:290-304  p1 = omega0, p2 = cSmago
 *   BOUNCEBACK BounceBack               core/dynamics.hh:783-791                        p1 = rho
 *   NODYNAMICS NoDynamics               core/dynamics.hh:968-972                        p1 = rho
 *   INCBGK     IncBGKdynamics           basicDynamics/isoThermalDynamics.hh:1405-1417   p1 = omega (rho0 = 1)   */
enum { PLB_B200_DYN_BGK = 0, PLB_B200_DYN_RLB = 1, PLB_B200_DYN_TRT = 2, PLB_B200_DYN_SMAGO = 3,
       PLB_B200_DYN_BOUNCEBACK = 4, PLB_B200_DYN_NODYNAMICS = 5, PLB_B200_DYN_INCBGK = 6 };
/* Boundary-condition families: LOCAL = createLocalBoundaryCondition3D (regularized,
 * boundaryCondition/boundaryCondition3D.hh:1162-1167), ZOUHE = createZouHeBoundaryCondition3D
 * (boundaryCondition/zouHeBoundary3D.hh:338-350). */
enum { PLB_B200_BC_LOCAL = 0, PLB_B200_BC_ZOUHE = 1, PLB_B200_BC_DYNAMICS_LOCAL = 2, PLB_B200_BC_EQUILIBRIUM = 3, PLB_B200_BC_INTERP = 4 };
/* INTERP = createInterpBoundaryCondition3D (boundaryCondition3D.hh:1049-1085, :1184-1188): Velocity / DensityDirichletBoundaryDynamics
 * on the plane cells + PlaneFdBoundaryFunctional3D (finiteDifferenceBoundaryProcessor3D.hh:68-117: velocity gradients by finite
 * differences of the neighbours' velocities, then regularize); outer edges and corners are those of LOCAL. */
/* DYNAMICS_LOCAL = createDynamicsBasedLocalBoundaryCondition3D, EQUILIBRIUM = createEquilibriumBoundaryCondition3D
 * (boundaryCondition3D.hh:1170-1181).  In this fork both managers (:712-747, :940-974) put the regularized dynamics on the plane
 * cells WITHOUT a data processor: the dynamics completes the populations at the beginning of collide()
 * (automaticPrepareCollision, core/dynamics.hh:470-475) and the cell holds raw populations between two steps; outer edges and
 * corners are those of LOCAL. */

typedef struct plb_b200_lattice plb_b200_lattice;

/* ---- life cycle ---------------------------------------------------------------------------------- */

/* MultiBlockLattice3D<T,Descriptor>(nx,ny,nz, backgroundDynamics) multiBlock/multiBlockLattice3D.hh:76-87.
 * q = 19 | 27 (D3Q19Descriptor / D3Q27Descriptor), precision_bytes = 8 | 4 (T = double | float).
 * The lattice lives on the current CUDA device as one x-slab. */
plb_b200_lattice* plb_b200_create(int q, int precision_bytes, int nx, int ny, int nz,
                                  int dynamics, double p1, double p2);

/* Same, decomposed into x-slabs (the reference's createXSlicedDistribution3D,
 * multiBlock/staticRepartitions3D.h:72-74): n_ranks processes (one GPU each) times slabs_per_rank
 * slabs in this process.  With n_ranks > 1, plb_b200_comm_init must follow before the first step. */
plb_b200_lattice* plb_b200_create_distributed(int q, int precision_bytes, int nx, int ny, int nz,
                                              int dynamics, double p1, double p2,
                                              int rank, int n_ranks, int slabs_per_rank);
/* A lattice that is stepped by the external-moment sweep of the fork's own application:
 *   generateMultiScalarField / generateMultiTensorField<3>(lattice, envelopeWidth)                 src/variables.hh:475-479
 *   integrateProcessingFunctional(new ExternalRhoJcollideAndStream3D, bbox, {lattice,rhoBar,j}, 0)  src/variables.hh:485
 *   integrateProcessingFunctional(new BoxRhoBarJfunctional3D,        bbox, {lattice,rhoBar,j}, 3)  src/variables.hh:486
 * That sweep works on the bulk ENLARGED by one cell (basicDynamics/dynamicsProcessor3D.hh:118-119): it collides the
 * reference's envelope layer -- with the zero moments found there -- and streams it into the rim of the lattice.  The engine
 * therefore allocates one extra layer on every non-periodic side (x slabs as usual) and shifts every box of this ABI; the
 * caller keeps the coordinates of its own nx x ny x nz lattice.  Periodicity is fixed here (the envelope of a periodic
 * direction is the periodic image, handled by index wrap).  The boundary condition is set up afterwards: its level-0
 * processors then run after the sweep, the order this engine implements.  Step with plb_b200_execute_internal_processors. */
plb_b200_lattice* plb_b200_create_enveloped(int q, int precision_bytes, int nx, int ny, int nz,
                                            int dynamics, double p1, double p2,
                                            int periodic_x, int periodic_y, int periodic_z,
                                            int rank, int n_ranks, int slabs_per_rank);
void plb_b200_destroy(plb_b200_lattice* L);
const char* plb_b200_last_error(void);
int plb_b200_abi_version(void);

/* Envelope exchange between ranks (replaces ParallelBlockCommunicator3D::communicate,
 * parallelism/parallelBlockCommunicator3D.cpp:251-297, MPI_Isend/Irecv of whole cells) by NCCL
 * send/recv of the face-crossing populations only.  Rank 0 obtains a 128-byte id, the host code
 * broadcasts it by whatever means it has (MPI_Bcast in Palabos, torch.distributed here), every rank
 * passes it in. */
int plb_b200_comm_unique_id(char id128[128]);
int plb_b200_comm_init(plb_b200_lattice* L, const char id128[128]);

/* Host-side description of the decomposition (no GPU needed): x range [x0, x1] of slab
 * `slab` out of `n_slabs` for a lattice of nx planes. */
int plb_b200_slab_range(int nx, int n_slabs, int slab, int* x0, int* x1);

/* CUDA stream on which every kernel of this lattice is launched (default: the legacy default
 * stream, which is also torch's default current stream). */
int plb_b200_set_stream(plb_b200_lattice* L, void* cuda_stream);

/* ---- set-up (the reference's wrappers) ------------------------------------------------------------ */

/* lattice.periodicity().toggle(direction, flag) multiBlock/multiBlock3D.h */
int plb_b200_set_periodicity(plb_b200_lattice* L, int px, int py, int pz);
/* lattice.toggleInternalStatistics(flag) multiBlock/multiBlock3D.cpp */
int plb_b200_toggle_internal_statistics(plb_b200_lattice* L, int on);
/* defineDynamics(lattice, box, dynamics) dataProcessors/dataInitializerWrapper3D.h:163-165 */
int plb_b200_define_dynamics_box(plb_b200_lattice* L, const int box[6], int dynamics, double p1, double p2);
/* defineDynamics(lattice, intMask, dynamics, flag) dataInitializerWrapper3D.h:194-200;
 * mask = int32 [nx][ny][nz] of the whole lattice */
int plb_b200_define_dynamics_mask(plb_b200_lattice* L, const int32_t* mask, int flag, int dynamics, double p1, double p2);
/* OnLatticeBoundaryCondition3D::addVelocityBoundary{0,1,2}{N,P} / addPressureBoundary..
 * boundaryCondition/boundaryCondition3D.h:40-455, instantiated by boundaryInstantiator3D.h:1016-1070 */
int plb_b200_add_velocity_boundary(plb_b200_lattice* L, int model, int direction, int orientation, const int box[6]);
int plb_b200_add_pressure_boundary(plb_b200_lattice* L, int model, int direction, int orientation, const int box[6]);
/* addExternalVelocityEdge{0,1,2}{NN,NP,PN,PP} / addExternalVelocityCorner{NNN..PPP}
 * boundaryCondition3D.hh:861-915 */
int plb_b200_add_external_velocity_edge(plb_b200_lattice* L, int model, int plane, int normal1, int normal2, const int box[6]);
int plb_b200_add_external_velocity_corner(plb_b200_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z);
/* addInternalVelocityEdge{0,1,2}{N,P}{N,P}(box, lattice) / addInternalVelocityCorner{N,P}{N,P}{N,P}(x, y, z, lattice)
 * boundaryCondition3D.h:123-206: the CONCAVE edges and corners of a wall (a solid block inside the fluid, a backward-facing
 * step): RegularizedVelocityInnerEdgeDynamics3D / RegularizedVelocityInnerCornerDynamics3D, regularizedBoundaryDynamics3D.hh:40-298.
 * Normals as for the external ones (from the fluid into the solid); completed by a processor under models 0 and 1, at
 * collision time under 2, 3 and 4, as in the reference's managers. */
int plb_b200_add_internal_velocity_edge(plb_b200_lattice* L, int model, int plane, int normal1, int normal2, const int box[6]);
int plb_b200_add_internal_velocity_corner(plb_b200_lattice* L, int model, int xn, int yn, int zn, int x, int y, int z);
/* OnLatticeBoundaryCondition3D::setVelocityConditionOnBlockBoundaries(lattice, block, applicationDomain)
 * boundaryCondition3D.hh:64-162 */
int plb_b200_set_velocity_condition_on_block_boundaries(plb_b200_lattice* L, int model, const int block[6], const int app[6]);
/* setBoundaryVelocity(lattice, box, u) / setBoundaryDensity(lattice, box, rho)
 * dataProcessors/dataInitializerFunctional3D.hh:830-845 */
int plb_b200_set_boundary_velocity(plb_b200_lattice* L, const int box[6], const double u[3]);
int plb_b200_set_boundary_density(plb_b200_lattice* L, const int box[6], double rho);
/* initializeAtEquilibrium(lattice, box, rho, u) dataInitializerWrapper3D.h; _field: the functional
 * overload, rho_u = double [cells of box in x,y,z order][rho,ux,uy,uz] */
int plb_b200_initialize_at_equilibrium(plb_b200_lattice* L, const int box[6], double rho, const double u[3]);
int plb_b200_initialize_at_equilibrium_field(plb_b200_lattice* L, const int box[6], const double* rho_u);
/* lattice.initialize(): MultiBlock3D::initialize -> executeInternalProcessors multiBlock/multiBlock3D.cpp:486-510 */
int plb_b200_initialize(plb_b200_lattice* L);

/* ---- data in / out (wire format above) ------------------------------------------------------------- */

/* transferToCoProcessors / CoProcessor3D::send coProcessors/coProcessorCommunication3D.hh:36-75.
 * `data` is HOST memory of exactly cells(box)*q*sizeof(T) bytes.  With several ranks each rank
 * passes / receives the part of the box it owns (its slabs), packed in x,y,z order. */
int plb_b200_upload(plb_b200_lattice* L, const int box[6], const void* data, size_t bytes);
int plb_b200_download(plb_b200_lattice* L, const int box[6], void* data, size_t bytes);
/* saveBinaryBlock(lattice, fName) / loadBinaryBlock(lattice, fName) io/serializerIO_3D.cpp:37-66: base64 of the data size
 * (64-bit) followed by base64 of the populations in the wire format above; files are interchangeable with the reference's
 * (checkpoint / restart).  One rank only. */
int plb_b200_save_binary_block(plb_b200_lattice* L, const char* path);
int plb_b200_load_binary_block(plb_b200_lattice* L, const char* path);
/* parallelIO::save(multiBlock, fName, dynamicContent = false) io/multiBlockWriter3D.cpp:226-238 and parallelIO::load
 * io/multiBlockReader3D.cpp:251-259: <base_path>.plb (the XML description: family, data type, descriptor, bounding box, one
 * component per x-slab with its byte offset) + <base_path>.dat (the populations of the slabs back to back, raw T, wire format
 * above -- byte-identical to the reference's file for the same blocks).  Files are interchangeable with the reference's: it
 * loads ours, and plb_b200_parallel_load reads a file the reference wrote with any block structure.  Static variables only: a
 * file saved with dynamicContent = true (serialized dynamics objects per cell) is refused -- the set-up lives in the program.
 * With several ranks every rank calls; each writes / reads the slabs it owns, rank 0 writes the .plb.
 * data_file_entry: the <File> entry of the .plb (NULL: <base_path>.dat); global_id: MultiBlock3D::getId() of the block. */
int plb_b200_parallel_save(plb_b200_lattice* L, const char* base_path, const char* data_file_entry, int global_id);
int plb_b200_parallel_load(plb_b200_lattice* L, const char* base_path);
/* bytes a rank must supply to upload/download for `box` (cells it owns inside the box * q * sizeof(T)) */
size_t plb_b200_local_bytes(plb_b200_lattice* L, const int box[6]);

/* ---- the hot path ----------------------------------------------------------------------------------- */

/* n x MultiBlockLattice3D::collideAndStream() multiBlock/multiBlockLattice3D.hh:318-360: fused
 * collide+stream sweep, halo exchange, level-0 boundary processors, statistics.  Asynchronous. */
int plb_b200_collide_and_stream(plb_b200_lattice* L, int n_steps);
/* n x { lattice.executeInternalProcessors(); lattice.incrementTime(); } viscosityTest/src/test.cpp:64-65 -- the time step of
 * a lattice made by plb_b200_create_enveloped: sweep with rhoBar and j read from the fields (Dynamics::collideExternal,
 * basicDynamics/dynamicsProcessor3D.hh:41-195), halo exchange, level-0 boundary processors, then Dynamics::computeRhoBarJ
 * of every cell of the lattice into the fields (dataProcessors/dataAnalysisFunctional3D.hh:528-548).  The internal
 * statistics are not evaluated, as in the reference's loop.  plb_b200_initialize on such a lattice is one of these steps
 * (MultiBlock3D::initialize = executeInternalProcessors, multiBlock/multiBlock3D.cpp:330-332). */
int plb_b200_execute_internal_processors(plb_b200_lattice* L, int n_steps);
/* applyProcessingFunctional(new BoxRhoBarJfunctional3D, box, rhoBarJarg) src/variables.hh:650 */
int plb_b200_apply_rhobar_j(plb_b200_lattice* L, const int box[6]);
/* the rhoBar (ScalarField3D) and j (TensorField3D<3>) fields, e.g. for the processors that force the momentum between two
 * steps: host memory, T [cells of box][rhoBar, jx, jy, jz] (the part of the box this rank owns, as plb_b200_upload) */
int plb_b200_upload_rhobar_j(plb_b200_lattice* L, const int box[6], const void* data, size_t bytes);
int plb_b200_download_rhobar_j(plb_b200_lattice* L, const int box[6], void* data, size_t bytes);
/* block until everything queued for this lattice has run */
int plb_b200_synchronize(plb_b200_lattice* L);

/* ---- results ------------------------------------------------------------------------------------------ */

/* internal statistics of the last step (core/latticeStatistics.h:32-34): out3 = {average rhoBar,
 * average uSqr, max uSqr}; getStoredAverageDensity = out3[0]+1, getStoredAverageEnergy = out3[1]/2,
 * getStoredMaxVelocity = sqrt(out3[2]) (multiBlockLattice3D.hh:531-548).  Reduced over all ranks. */
int plb_b200_get_stored_statistics(plb_b200_lattice* L, double out3[3], uint64_t* n_cells);
/* computeAverageEnergy / computeAverageDensity dataProcessors/dataAnalysisWrapper3D.hh:1252-1262, :1226-1230 */
int plb_b200_compute_average_energy(plb_b200_lattice* L, const int box[6], double* out);
int plb_b200_compute_average_density(plb_b200_lattice* L, const int box[6], double* out);
/* computeDensity + computeVelocity fields: rho_u = T [cells][rho,ux,uy,uz] (host memory) */
int plb_b200_compute_density_velocity(plb_b200_lattice* L, const int box[6], void* rho_u, size_t bytes);

/* ---- introspection ------------------------------------------------------------------------------------- */

/* kernels launched by this lattice since creation (bench.py's gpu_launches) */
uint64_t plb_b200_launch_count(plb_b200_lattice* L);
/* device bytes held by this lattice */
size_t plb_b200_device_bytes(plb_b200_lattice* L);
/* number of boundary stages per step, and items in them */
int plb_b200_boundary_stages(plb_b200_lattice* L, int* n_stages, int64_t* n_items);

/* CUDA-event timing of the sweep kernel alone (the reference's profiler timer "collStream",
 * atomicBlock/blockLattice3D.hh:257-258,311): switch on, step, read total milliseconds and launches. */
int plb_b200_profile_sweeps(plb_b200_lattice* L, int on);
int plb_b200_profile_read(plb_b200_lattice* L, double* total_ms, int* n_launches);
/* how the envelopes between x-slabs are refreshed in this lattice (a static string; replaces the choice of
 * BlockCommunicator3D, multiBlock/serialBlockCommunicator3D.cpp / parallelism/parallelBlockCommunicator3D.cpp):
 * direct stores of the sweep's rim cells into the neighbouring slab (CUDA IPC over NVLink between ranks), or -- with
 * PLB_B200_HALO=nccl in the environment, or where peer mapping is impossible -- ghost planes + ncclSend/ncclRecv. */
const char* plb_b200_halo_transport(plb_b200_lattice* L);

/* ---- geometry side of the path: STL -> voxel flags --------------------------------- */

/* voxelize(mesh, domain, borderWidth) offLattice/voxelizer.hh:107-142 + detectBorderLine :640-690, on the device.
 * triangles: double [n][3 vertices][x,y,z] in lattice units (what TriangularSurfaceMesh hands to the voxelizer);
 * flags_out: int32 [nx][ny][nz] (host), voxelFlag values offLattice/voxelizer.h:38-48: 1 outside, 2 outerBorder,
 * 3 inside, 4 innerBorder.  kernel_ms (optional): device time of the kernels.  The matrix is what
 * defineDynamics(lattice, mask, flag, dynamics) (plb_b200_define_dynamics_mask) takes. */
int plb_b200_voxelize(const double* triangles, int64_t n_triangles, int nx, int ny, int nz, int border_width,
                      int32_t* flags_out, double* kernel_ms);
/* TriangleSet<T>(fileName, FLT) offLattice/triangleSet.hh:130-350: ASCII (numbers read as double) or binary STL ->
 * double [n][3][3] (malloc'ed, release with plb_b200_stl_free); degenerate facets are dropped and vertices 1, 2 swapped
 * where the facet normal of the file says so (checkForDegenerateTrianglesAndFixOrientationNoAbort :497-548) */
int plb_b200_stl_read(const char* path, double** soup, int64_t* n_triangles);
void plb_b200_stl_free(double* soup);
/* DEFscaledMesh(triangleSet, resolution, referenceDirection, margin, extraLayer) offLattice/triangleBoundary3D.hh:255-301
 * followed by TriangularSurfaceMesh::inflate(inflate) triangularSurfaceMesh.hh:1419-1425 (reference default 1e-3):
 * triangles_out double [n][3][3] in lattice units, dims = size of the voxel matrix (voxelizer.hh:84-104),
 * dx / location = physical size of a cell and position of lattice node (0,0,0). */
int plb_b200_mesh_to_lattice(const double* soup, int64_t n_triangles, int resolution, int reference_direction, int margin,
                             int extra_layer, double inflate, double* triangles_out, int dims[3], double* dx, double location[3]);

/* ---- the reference's own plug-in surface: plb::CoProcessor3D<T> (coProcessor3D.h:59-116) ------------- */
/* D3Q19 BGK only, domain = whole block incl. envelope, result undefined in the outer 1-cell layer.
 * These four are what global::defaultCoProcessor3D<T>() (coProcessor3D.h:179-186) should forward to. */
int plb_b200_coproc_add_domain(int precision_bytes, int nx, int ny, int nz, double omega, int* domain_handle);
int plb_b200_coproc_send(int domain_handle, const int sub_domain[6], const char* data, size_t bytes);
int plb_b200_coproc_receive(int domain_handle, const int sub_domain[6], char* data, size_t bytes);
int plb_b200_coproc_collide_and_stream(int domain_handle);
int plb_b200_coproc_remove_domain(int domain_handle);

#ifdef __cplusplus
}
#endif
#endif
