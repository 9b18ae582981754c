"""Parity cases: small versions of BASELINE.json's configs plus edge cases.

Each builder returns a tests.scenario.Scenario.  oracle/gen_golden.py replays them on the real
reference (oracle/_ref/ref_driver) and commits the outcome under tests/golden/; the tests replay
them on the oracle and on the CUDA engine.  Test infrastructure only.
"""
from __future__ import annotations

import numpy as np

from scenario import Scenario


def cavity(nx, ny, nz, steps, dyn="BGK", q=19, prec="f64", blocks=(1, 1, 1), every=None, model="local",
           dumps=True):
    """3D lid-driven cavity of jlabos/src/testing/test_cylinder.py:100-119 (BASELINE configs[0], [1]):
    uMax=0.02, Re=100, nu=uMax*ny/Re, omega=1/(3nu+0.5); regularized velocity BC on all faces,
    edges, corners; lid z=nz-1 interior moves with (uMax,0,0); rho=1, u=0 initially."""
    u_max, re = 0.02, 100.0
    nu = u_max * ny / re
    omega = 1.0 / (3 * nu + 0.5)
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, 0.14 if dyn == "SMAGO" else 0.0, blocks)
    sc.bc_box(model)
    sc.set_velocity(sc.box, (0, 0, 0))
    sc.set_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (u_max, 0, 0))
    sc.init_eq(sc.box, 1.0, (0, 0, 0))
    sc.initialize()
    if dumps:
        sc.dump("f_init")
    done = 0
    if every:
        while done + every <= steps:
            sc.run(every).stats()
            done += every
    if steps > done:
        sc.run(steps - done)
    sc.stats()
    if dumps:
        sc.dump("f_end").dump_rho_u("rho_u_end")
    return sc


def poiseuille_profile(nx, ny, nz, u_max, axis_wall=1, perturb=0.0):
    """rho=1, parabolic u_x(y) between walls at y=0 and y=ny-1, optional deterministic sinusoidal
    perturbation."""
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    yy = (y - 0.5) / (ny - 2.0)
    ux = 4.0 * u_max * yy * (1.0 - yy)
    ux[(y == 0) | (y == ny - 1)] = 0.0
    ru = np.zeros((nx, ny, nz, 4))
    ru[..., 0] = 1.0
    ru[..., 1] = ux
    if perturb:
        ru[..., 1] += perturb * np.sin(2 * np.pi * (y + 0.5) / ny) * np.cos(2 * np.pi * (z + 0.25) / nz)
        ru[..., 2] += perturb * np.sin(2 * np.pi * (x + 0.5) / nx) * np.sin(2 * np.pi * (z + 0.75) / nz)
        ru[..., 3] += perturb * np.cos(2 * np.pi * (x + 0.25) / nx) * np.sin(2 * np.pi * (y + 0.5) / ny)
        ru[..., 0] += 0.1 * perturb * np.cos(2 * np.pi * x / nx)
    return ru


def channel(nx, ny, nz, steps, dyn="TRT", omega=1.2, q=19, prec="f64", cs=0.0, perturb=0.0, u_max=0.05,
            blocks=(1, 1, 1), every=None):
    """Periodic channel (BASELINE configs[2] with TRT; configs[4] with D3Q27 Smagorinsky): periodic in
    x and z, BounceBack walls at y=0 and y=ny-1, parabolic initial profile."""
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, cs, blocks)
    sc.periodic(1, 0, 1)
    sc.dynamics((0, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
    sc.dynamics((0, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
    sc.init_eq_field(sc.box, poiseuille_profile(nx, ny, nz, u_max, perturb=perturb))
    sc.initialize()
    sc.dump("f_init")
    done = 0
    if every:
        while done + every <= steps:
            sc.run(every).stats()
            done += every
    if steps > done:
        sc.run(steps - done)
    sc.stats().dump("f_end").dump_rho_u("rho_u_end")
    return sc


def sphere_mask(nx, ny, nz, cx, cy, cz, r):
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    return (((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) <= r * r).astype(np.int32)


def obstacle(nx, ny, nz, steps, dyn="BGK", omega=1.5, model="zouhe", q=19, prec="f64", u_in=0.03, mask=None,
             blocks=(1, 1, 1)):
    """Flow past a voxelised obstacle (BASELINE configs[3]): velocity inlet x=0 and pressure outlet
    x=nx-1 (Zou-He or regularized), BounceBack on the four side walls and on the flagged voxels."""
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, 0.0, blocks)
    sc.bc_vel(model, 0, -1, (0, 0, 1, ny - 2, 1, nz - 2))
    sc.bc_pres(model, 0, +1, (nx - 1, nx - 1, 1, ny - 2, 1, nz - 2))
    sc.dynamics((0, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
    sc.dynamics((0, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
    sc.dynamics((0, nx - 1, 0, ny - 1, 0, 0), "BB", 1.0)
    sc.dynamics((0, nx - 1, 0, ny - 1, nz - 1, nz - 1), "BB", 1.0)
    if mask is None:
        mask = sphere_mask(nx, ny, nz, nx // 3, ny / 2 - 0.5, nz / 2 - 0.5, min(ny, nz) / 5.0)
    sc.dynamics_mask(mask, 1, "BB", 1.0)
    sc.set_velocity(sc.box, (u_in, 0, 0))
    sc.set_density(sc.box, 1.0)
    sc.init_eq(sc.box, 1.0, (u_in, 0, 0))
    sc.initialize()
    sc.dump("f_init")
    sc.run(steps).stats().dump("f_end").dump_rho_u("rho_u_end")
    return sc


def periodic_box(nx, ny, nz, steps, dyn, omega, q=19, prec="f64", cs=0.0, perturb=0.02, blocks=(1, 1, 1)):
    """Fully periodic box with a deterministic perturbation; optional region of different omega."""
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, cs, blocks)
    sc.periodic(1, 1, 1)
    ru = poiseuille_profile(nx, ny, nz, 0.0, perturb=perturb)
    ru[..., 1] += 0.03
    sc.init_eq_field(sc.box, ru)
    sc.initialize()
    sc.run(steps).stats().dump("f_end").dump_rho_u("rho_u_end")
    return sc


def mixed_omega(nx, ny, nz, steps):
    """BGK with a sub-box at a different omega, a NoDynamics block and half-open boundaries: exercises the
    dynamics table, the on-site rule at non-periodic domain edges and partial periodicity."""
    sc = Scenario(19, "f64", nx, ny, nz, "BGK", 1.3)
    sc.periodic(0, 1, 0)
    sc.dynamics((2, nx // 2, 1, ny - 2, 2, nz - 3), "BGK", 1.7)
    sc.dynamics((nx - 3, nx - 2, 2, 3, 1, 2), "NODYN", 1.0)
    sc.dynamics((1, 2, 1, 2, nz - 2, nz - 2), "RLB", 1.1)
    ru = poiseuille_profile(nx, ny, nz, 0.0, perturb=0.02)
    sc.init_eq_field(sc.box, ru)
    sc.initialize()
    sc.run(steps).stats().dump("f_end").dump_rho_u("rho_u_end")
    return sc


def inlet_outlet_local(nx, ny, nz, steps):
    """Regularized velocity inlet + regularized pressure outlet + velocity side walls with their outer
    edges/corners restricted to a sub-domain (applicationDomain != block)."""
    sc = Scenario(19, "f64", nx, ny, nz, "BGK", 1.6)
    sc.bc_box("local", sc.box, (0, nx - 2, 0, ny - 1, 0, nz - 1))          # everything but the x=nx-1 face
    sc.bc_pres("local", 0, +1, (nx - 1, nx - 1, 1, ny - 2, 1, nz - 2))
    sc.dynamics((nx - 1, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
    sc.dynamics((nx - 1, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
    sc.dynamics((nx - 1, nx - 1, 0, ny - 1, 0, 0), "BB", 1.0)
    sc.dynamics((nx - 1, nx - 1, 0, ny - 1, nz - 1, nz - 1), "BB", 1.0)
    sc.set_velocity(sc.box, (0, 0, 0))
    sc.set_velocity((0, 0, 1, ny - 2, 1, nz - 2), (0.02, 0.001, -0.002))
    sc.set_density(sc.box, 1.01)
    sc.init_eq(sc.box, 1.0, (0.01, 0, 0))
    sc.initialize()
    sc.dump("f_init")
    sc.run(steps).stats().dump("f_end").dump_rho_u("rho_u_end")
    return sc


# ---- the external-moment sweep: the time loop of the fork's own application -------------------
# src/variables.hh:475-486 integrates ExternalRhoJcollideAndStream3D at level 0 and BoxRhoBarJfunctional3D at level 3 on
# {lattice, rhoBar, j}; :650 fills the fields once; viscosityTest/src/test.cpp:64-65 steps with executeInternalProcessors() +
# incrementTime().  ext_add_j stands for the processors that force the momentum field between two steps.


def ext_finish(sc, steps, add_j):
    nx, ny, nz = sc.nx, sc.ny, sc.nz
    sc.ext_apply()
    sc.dump_rhobar_j("rj_init")
    first = steps // 2
    sc.ext_exec(first)
    if add_j:
        sc.ext_add_j((1, max(1, nx // 2), 1, ny - 2, max(1, nz // 3), nz - 2), (1e-3, -5e-4, 2.5e-4))
    sc.ext_exec(steps - first)
    return sc.stats().dump("f_end").dump_rhobar_j("rj_end").dump_rho_u("rho_u_end")


def ext_cavity(nx, ny, nz, steps, dyn="INCBGK", q=19, prec="f64", model="local", add_j=True, initialize=False):
    u_max, re = 0.02, 100.0
    omega = 1.0 / (3 * u_max * ny / re + 0.5)
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, 0.14 if dyn == "SMAGO" else 0.0)
    sc.ext_integrate()
    sc.bc_box(model)
    sc.set_velocity(sc.box, (0, 0, 0))
    sc.set_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (u_max, 0, 0))
    sc.init_eq(sc.box, 1.0, (0, 0, 0))
    if initialize:
        sc.ext_apply()
        sc.initialize()        # MultiBlock3D::initialize = executeInternalProcessors: a whole step in this mode
    return ext_finish(sc, steps, add_j)


def ext_channel(nx, ny, nz, steps, dyn="TRT", omega=1.2, q=19, prec="f64", cs=0.0, perturb=0.01, add_j=True):
    """BounceBack walls ON the rim of the lattice: what streams in from the reference's envelope layer (collided with
    the zero moments found there) matters."""
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, cs)
    sc.periodic(1, 0, 1)
    sc.ext_integrate()
    sc.dynamics((0, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
    sc.dynamics((0, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
    sc.dynamics((2, 3, 2, 3, 1, 2), "NODYN", 1.0)
    sc.init_eq_field(sc.box, poiseuille_profile(nx, ny, nz, 0.05, perturb=perturb))
    return ext_finish(sc, steps, add_j)


def ext_open(nx, ny, nz, steps, dyn="BGK", omega=1.5, q=19, prec="f64"):
    """No boundary condition at all on a non-periodic lattice: bulk cells on the rim exchange populations with the
    envelope layer every step."""
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, 0.0)
    sc.periodic(0, 1, 0)
    sc.ext_integrate()
    sc.dynamics((2, nx // 2, 1, ny - 2, 2, nz - 3), dyn, 1.7)
    ru = poiseuille_profile(nx, ny, nz, 0.0, perturb=0.02)
    ru[..., 1] += 0.02
    sc.init_eq_field(sc.box, ru)
    return ext_finish(sc, steps, True)


def ext_obstacle(nx, ny, nz, steps, dyn="INCBGK", omega=1.6, model="zouhe"):
    sc = Scenario(19, "f64", nx, ny, nz, dyn, omega, 0.0)
    sc.ext_integrate()
    sc.bc_vel(model, 0, -1, (0, 0, 1, ny - 2, 1, nz - 2))
    sc.bc_pres(model, 0, +1, (nx - 1, nx - 1, 1, ny - 2, 1, nz - 2))
    sc.dynamics((0, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
    sc.dynamics((0, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
    sc.dynamics((0, nx - 1, 0, ny - 1, 0, 0), "BB", 1.0)
    sc.dynamics((0, nx - 1, 0, ny - 1, nz - 1, nz - 1), "BB", 1.0)
    sc.dynamics_mask(sphere_mask(nx, ny, nz, nx // 3, ny / 2 - 0.5, nz / 2 - 0.5, min(ny, nz) / 5.0), 1, "BB", 1.0)
    sc.set_velocity(sc.box, (0.03, 0, 0))
    sc.set_density(sc.box, 1.0)
    sc.init_eq(sc.box, 1.0, (0.03, 0, 0))
    return ext_finish(sc, steps, False)


def ext_periodic(nx, ny, nz, steps, dyn, omega, q=19, prec="f64", cs=0.0):
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, cs)
    sc.periodic(1, 1, 1)
    sc.ext_integrate()
    ru = poiseuille_profile(nx, ny, nz, 0.0, perturb=0.02)
    ru[..., 1] += 0.03
    sc.init_eq_field(sc.box, ru)
    return ext_finish(sc, steps, True)


def ext_interleaved(nx=10, ny=8, nz=9):
    """Everything a caller can do between two steps of the external-moment loop, interleaved with steps: read the fields,
    overwrite the populations, change dynamics and boundary values, recompute the fields on a part of the lattice, force j.
    (On the engine: every transition between 'fields implied by the populations' and 'fields are data of their own'.)"""
    sc = Scenario(19, "f64", nx, ny, nz, "BGK", 1.45, 0.0)
    sc.periodic(0, 0, 1)
    sc.ext_integrate()
    sc.bc_vel("local", 0, -1, (0, 0, 1, ny - 2, 0, nz - 1))
    sc.bc_pres("zouhe", 0, +1, (nx - 1, nx - 1, 1, ny - 2, 0, nz - 1))
    sc.dynamics((0, nx - 1, 0, 0, 0, nz - 1), "BB", 1.0)
    sc.dynamics((0, nx - 1, ny - 1, ny - 1, 0, nz - 1), "BB", 1.0)
    sc.set_velocity(sc.box, (0.02, 0, 0))
    sc.set_density(sc.box, 1.0)
    ru = poiseuille_profile(nx, ny, nz, 0.02, perturb=0.005)
    sc.init_eq_field(sc.box, ru)
    sc.ext_exec(2)                                  # zero fields: nobody has applied BoxRhoBarJ yet
    sc.dump_rhobar_j("rj_a")
    sc.ext_exec(3)
    k = np.arange(nx * ny * nz * 19, dtype=np.float64).reshape(nx, ny, nz, 19)
    sc.upload(1e-3 * np.sin(0.37 * k) * (1.0 / 19))  # populations replaced, fields keep the old moments
    sc.ext_exec(2)
    sc.dynamics((3, 5, 2, 4, 1, 3), "RLB", 1.2)
    sc.dynamics((6, 7, 3, 4, 4, 6), "BB", 1.0)
    sc.ext_exec(2)
    sc.set_velocity((0, 0, 1, ny - 2, 0, nz - 1), (0.03, 0.001, 0))
    sc.ext_exec(2)
    sc.ext_add_j((2, 6, 1, 5, 2, 7), (5e-4, 0, -2e-4))
    sc.ext_apply((4, nx - 1, 0, ny - 1, 0, 4))       # BoxRhoBarJ on a part: undoes the forcing there only
    sc.dump_rhobar_j("rj_b")
    sc.ext_exec(1)
    sc.init_eq((1, 3, 1, 3, 1, 3), 1.02, (0.01, 0.0, -0.01))
    sc.ext_exec(2)
    sc.ext_apply()
    sc.ext_exec(1)
    return sc.stats().dump("f_end").dump_rhobar_j("rj_end").dump_rho_u("rho_u_end")


def inlet_outlet_mixed(nx, ny, nz, steps):
    """Wrapped (processor-completed) and dynamics-based (collision-time) plane boundaries in one lattice: a regularized
    velocity inlet completed by its processor, a dynamics-based pressure outlet, dynamics-based side walls with their outer
    edges; BounceBack where nothing else sits."""
    sc = Scenario(19, "f64", nx, ny, nz, "BGK", 1.55)
    sc.bc_vel("local", 0, -1, (0, 0, 1, ny - 2, 1, nz - 2))
    sc.bc_pres("dynlocal", 0, +1, (nx - 1, nx - 1, 1, ny - 2, 1, nz - 2))
    sc.bc_vel("dynlocal", 1, -1, (1, nx - 2, 0, 0, 1, nz - 2))
    sc.bc_vel("equilibrium", 1, +1, (1, nx - 2, ny - 1, ny - 1, 1, nz - 2))
    sc.bc_vel("zouhe", 2, -1, (1, nx - 2, 1, ny - 2, 0, 0))
    sc.bc_vel("dynlocal", 2, +1, (1, nx - 2, 1, ny - 2, nz - 1, nz - 1))
    sc.bc_edge("dynlocal", 0, -1, -1, (1, nx - 2, 0, 0, 0, 0))
    sc.bc_edge("local", 0, +1, +1, (1, nx - 2, ny - 1, ny - 1, nz - 1, nz - 1))
    for wall in ((0, 0, 0, ny - 1, 0, 0), (0, 0, 0, ny - 1, nz - 1, nz - 1), (0, 0, 0, 0, 0, nz - 1), (0, 0, ny - 1, ny - 1, 0, nz - 1),
                 (nx - 1, nx - 1, 0, ny - 1, 0, 0), (nx - 1, nx - 1, 0, ny - 1, nz - 1, nz - 1), (nx - 1, nx - 1, 0, 0, 0, nz - 1),
                 (nx - 1, nx - 1, ny - 1, ny - 1, 0, nz - 1), (1, nx - 2, 0, 0, nz - 1, nz - 1), (1, nx - 2, ny - 1, ny - 1, 0, 0)):
        sc.dynamics(wall, "BB", 1.0)
    sc.set_velocity(sc.box, (0, 0, 0))
    sc.set_velocity((0, 0, 1, ny - 2, 1, nz - 2), (0.02, 0.001, -0.002))
    sc.set_density(sc.box, 1.005)
    sc.init_eq(sc.box, 1.0, (0.01, 0, 0))
    sc.initialize()
    sc.dump("f_init")
    sc.run(steps // 2).stats()
    sc.set_velocity((0, 0, 1, ny - 2, 1, nz - 2), (0.025, 0.0, 0.001))
    sc.run(steps - steps // 2).stats().dump("f_end").dump_rho_u("rho_u_end")
    return sc


def cavity_eraser(nx, ny, nz, steps, model="local", q=19, prec="f64"):
    """defineDynamics used as an "eraser tool" AFTER the boundary set-up (wrappedLocalBoundaryProcessor3D.hh:51-58): BounceBack,
    NoDynamics and plain bulk dynamics replace the boundary composites on parts of planes, edges and corners, by box and by
    mask; the data processors stay and meet cells without a composite.  (Engine: boundary records die, ADVICE r1 high.)"""
    u_max, re = 0.02, 100.0
    omega = 1.0 / (3 * u_max * ny / re + 0.5)
    sc = Scenario(q, prec, nx, ny, nz, "BGK", omega, 0.0)
    sc.bc_box(model)
    sc.set_velocity(sc.box, (0, 0, 0))
    sc.set_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (u_max, 0, 0))
    sc.dynamics((0, 0, 2, ny - 3, 1, nz // 2), "BB", 1.0)                  # part of the x = 0 plane
    sc.dynamics((2, 4, 3, 4, nz - 1, nz - 1), "BB", 1.0)                   # a patch of the lid
    sc.dynamics((1, nx // 2, 0, 0, 0, 0), "BB", 1.0)                       # half an edge
    sc.dynamics((nx - 1, nx - 1, ny - 1, ny - 1, nz - 1, nz - 1), "BB", 1.0)   # a corner
    sc.dynamics((0, 0, 0, 0, 0, 0), "NODYN", 1.0)                          # another corner
    sc.dynamics((nx - 1, nx - 1, 1, 3, 2, 4), "BGK", 1.3)                  # bulk dynamics on plane cells: the composite goes
    sc.dynamics((nx - 1, nx - 1, 0, 0, 2, 3), "RLB", 1.2)                  # ... and on edge cells
    mask = np.zeros((nx, ny, nz), dtype=np.int32)
    mask[3:6, ny - 1, 2:5] = 7                                             # patch of the y = ny-1 plane
    mask[nx - 2, ny - 1, 0] = 7                                            # an edge cell
    mask[5, 5, 3] = 7                                                      # and a bulk cell
    sc.dynamics_mask(mask, 7, "BB", 1.0)
    sc.init_eq(sc.box, 1.0, (0, 0, 0))
    sc.initialize()
    sc.dump("f_init")
    sc.run(steps // 2).stats()
    sc.set_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (u_max, 0.001, 0))
    sc.run(steps - steps // 2).stats().dump("f_end").dump_rho_u("rho_u_end")
    return sc


# name -> (builder, how to store the reference outcome: "full" dumps, or "probe" = sha256 + sampled cells)
def cavity_block(nx, ny, nz, steps, model="local", dyn="BGK", q=19, prec="f64", outer="local", u_block=(0.0, 0.0, 0.0)):
    """Lid-driven cavity with a solid block inside: the block's faces are velocity planes seen from the fluid around it, its
    twelve edges and eight corners are the CONCAVE (internal) edges and corners of the boundary condition
    (addInternalVelocityEdge* / addInternalVelocityCorner*, RegularizedVelocityInnerEdge/CornerDynamics3D), its inside is
    BounceBack.  Normals point from the fluid into the solid, as for every wall."""
    u_max, re = 0.02, 100.0
    omega = 1.0 / (3 * u_max * ny / re + 0.5)
    sc = Scenario(q, prec, nx, ny, nz, dyn, omega, 0.14 if dyn == "SMAGO" else 0.0)
    sc.bc_box(outer)
    x0, x1, y0, y1, z0, z1 = nx // 3, nx // 3 + 3, ny // 3, ny // 3 + 3, nz // 3, nz // 3 + 2
    lo, hi = (x0, y0, z0), (x1, y1, z1)
    sc.dynamics((x0 + 1, x1 - 1, y0 + 1, y1 - 1, z0 + 1, z1 - 1), "BB", 1.0)
    for d in range(3):                                   # faces: the block's low face has its solid at +d, i.e. orientation +1
        for ori, at in ((+1, lo[d]), (-1, hi[d])):
            b = [lo[0] + 1, hi[0] - 1, lo[1] + 1, hi[1] - 1, lo[2] + 1, hi[2] - 1]
            b[2 * d] = b[2 * d + 1] = at
            sc.bc_vel(model, d, ori, tuple(b))
    for plane in range(3):                               # edges parallel to `plane`: normals along (plane+1)%3 and (plane+2)%3
        d1, d2 = (plane + 1) % 3, (plane + 2) % 3
        for n1, a1 in ((+1, lo[d1]), (-1, hi[d1])):
            for n2, a2 in ((+1, lo[d2]), (-1, hi[d2])):
                b = [0] * 6
                b[2 * plane], b[2 * plane + 1] = lo[plane] + 1, hi[plane] - 1
                b[2 * d1] = b[2 * d1 + 1] = a1
                b[2 * d2] = b[2 * d2 + 1] = a2
                sc.bc_iedge(model, plane, n1, n2, tuple(b))
    for xn, x in ((+1, x0), (-1, x1)):
        for yn, y in ((+1, y0), (-1, y1)):
            for zn, z in ((+1, z0), (-1, z1)):
                sc.bc_icorner(model, xn, yn, zn, x, y, z)
    sc.set_velocity(sc.box, (0, 0, 0))
    sc.set_velocity((1, nx - 2, 1, ny - 2, nz - 1, nz - 1), (u_max, 0, 0))
    sc.set_velocity((x0, x1, y0, y1, z0, z1), u_block)
    sc.init_eq(sc.box, 1.0, (0, 0, 0))
    sc.initialize()
    sc.dump("f_init")
    sc.run(steps // 2).stats()
    sc.set_velocity((x0, x1, y0, y1, z0, z1), tuple(0.5 * v for v in u_block))
    sc.run(steps - steps // 2).stats()
    sc.dump("f_end").dump_rho_u("rho_u_end")
    return sc


CASES = {
    # configs[0]-style cavity, all dynamics on the path
    "cavity_bgk_10x9x8":        (lambda: cavity(10, 9, 8, 60, every=20), "full"),
    "cavity_rlb_10x9x8":        (lambda: cavity(10, 9, 8, 40, dyn="RLB"), "full"),
    "cavity_trt_9x8x10":        (lambda: cavity(9, 8, 10, 40, dyn="TRT"), "full"),
    "cavity_smago_8x9x10":      (lambda: cavity(8, 9, 10, 40, dyn="SMAGO"), "full"),
    "cavity_bgk_f32_10x9x8":    (lambda: cavity(10, 9, 8, 40, prec="f32"), "full"),
    "cavity_bgk_q27_8x8x9":     (lambda: cavity(8, 8, 9, 30, q=27), "full"),
    "cavity_zouhe_10x9x8":      (lambda: cavity(10, 9, 8, 40, model="zouhe"), "full"),
    # concave (internal) edges and corners: a solid block inside the cavity
    "block_local_13x12x11":     (lambda: cavity_block(13, 12, 11, 30, u_block=(0.004, -0.002, 0.003)), "full"),
    "block_local_q27_12x12x11": (lambda: cavity_block(12, 12, 11, 24, q=27, u_block=(0.003, 0.002, -0.002)), "full"),
    "block_zouhe_f32_13x11x12": (lambda: cavity_block(13, 11, 12, 24, model="zouhe", outer="zouhe", prec="f32"), "full"),
    "block_dynlocal_rlb_12x12x12": (lambda: cavity_block(12, 12, 12, 24, model="dynlocal", dyn="RLB", u_block=(0.0, 0.003, 0.0)), "full"),
    "block_incbgk_12x13x11":    (lambda: cavity_block(12, 13, 11, 24, dyn="INCBGK", u_block=(0.002, 0.0, 0.001)), "full"),
    # the reference's own known-answer sizes (test_cylinder.py:85-97), 30 steps
    "cavity_bgk_20":            (lambda: cavity(20, 20, 20, 30), "probe"),
    "cavity_bgk_39":            (lambda: cavity(39, 39, 39, 30), "probe"),
    "cavity_bgk_58":            (lambda: cavity(58, 58, 58, 30), "probe"),
    "cavity_bgk_20x31x42":      (lambda: cavity(20, 31, 42, 30), "probe"),
    "cavity_bgk_f32_20x31x42":  (lambda: cavity(20, 31, 42, 30, prec="f32"), "probe"),
    # decomposition invariance of the reference itself: same case, 8 blocks
    "cavity_bgk_12_blocks222":  (lambda: cavity(12, 12, 12, 30, blocks=(2, 2, 2)), "full"),
    "cavity_bgk_12":            (lambda: cavity(12, 12, 12, 30), "full"),
    # configs[0] proper: 64^3, 1000 steps (the 1e-10 pin of north_star)
    "cavity_bgk_64_1000":       (lambda: cavity(64, 64, 64, 1000, every=250, dumps=True), "probe"),
    # configs[2]: periodic TRT channel
    "channel_trt_16x12x10":     (lambda: channel(16, 12, 10, 50, every=25), "full"),
    "channel_trt_blocks":       (lambda: channel(16, 12, 10, 50, blocks=(4, 1, 1)), "full"),
    "channel_trt_q27_8x8x8":    (lambda: channel(8, 8, 8, 30, q=27), "full"),
    # configs[3]: obstacle with Zou-He inlet/outlet
    "obstacle_zouhe_24x12x10":  (lambda: obstacle(24, 12, 10, 50), "full"),
    "obstacle_local_rlb":       (lambda: obstacle(20, 10, 9, 40, dyn="RLB", omega=1.4, model="local"), "full"),
    # configs[4]: D3Q27 Smagorinsky channel, fp64 and fp32
    "channel_smago_q27_f64":    (lambda: channel(12, 10, 8, 50, dyn="SMAGO", omega=1.9, q=27, cs=0.14, perturb=0.01), "full"),
    "channel_smago_q27_f32":    (lambda: channel(12, 10, 8, 50, dyn="SMAGO", omega=1.9, q=27, cs=0.14, perturb=0.01, prec="f32"), "full"),
    "periodic_smago_q19_f32":   (lambda: periodic_box(9, 8, 7, 30, "SMAGO", 1.8, cs=0.17, prec="f32"), "full"),
    "periodic_bgk_q27_f32":     (lambda: periodic_box(8, 7, 9, 30, "BGK", 1.5, q=27, prec="f32"), "full"),
    "periodic_rlb_q27":         (lambda: periodic_box(8, 7, 9, 30, "RLB", 1.5, q=27), "full"),
    "periodic_trt_f32":         (lambda: periodic_box(8, 7, 9, 30, "TRT", 1.5, prec="f32"), "full"),
    # IncBGKdynamics (the dynamics of the fork's own application, viscosityTest): bulk, under regularized / Zou-He boundaries,
    # around BounceBack voxels, D3Q27 and fp32
    "cavity_incbgk_10x9x8":     (lambda: cavity(10, 9, 8, 40, dyn="INCBGK"), "full"),
    "cavity_incbgk_q27_f32":    (lambda: cavity(8, 8, 9, 30, dyn="INCBGK", q=27, prec="f32"), "full"),
    "obstacle_incbgk_zouhe":    (lambda: obstacle(20, 10, 9, 40, dyn="INCBGK", omega=1.6), "full"),
    "periodic_incbgk_q27":      (lambda: periodic_box(8, 7, 9, 30, "INCBGK", 1.7, q=27), "full"),
    # the external-moment sweep (ExternalRhoJcollideAndStream3D + BoxRhoBarJfunctional3D): the fork's own time loop
    "ext_cavity_incbgk":        (lambda: ext_cavity(10, 9, 8, 40), "full"),
    "ext_cavity_bgk_init":      (lambda: ext_cavity(9, 8, 10, 30, dyn="BGK", initialize=True), "full"),
    "ext_cavity_rlb_zouhe":     (lambda: ext_cavity(8, 9, 10, 30, dyn="RLB", model="zouhe"), "full"),
    "ext_cavity_smago_q27_f32": (lambda: ext_cavity(8, 8, 9, 30, dyn="SMAGO", q=27, prec="f32"), "full"),
    "ext_channel_trt":          (lambda: ext_channel(12, 9, 8, 40), "full"),
    "ext_channel_smago_q27":    (lambda: ext_channel(10, 8, 8, 30, dyn="SMAGO", omega=1.9, q=27, cs=0.14), "full"),
    "ext_channel_bgk_f32":      (lambda: ext_channel(10, 8, 9, 30, dyn="BGK", omega=1.4, prec="f32"), "full"),
    "ext_open_bgk":             (lambda: ext_open(11, 8, 9, 30), "full"),
    "ext_open_incbgk_q27":      (lambda: ext_open(9, 8, 8, 30, dyn="INCBGK", q=27), "full"),
    "ext_obstacle_incbgk":      (lambda: ext_obstacle(20, 10, 9, 40), "full"),
    "ext_periodic_trt_q27":     (lambda: ext_periodic(8, 7, 9, 30, "TRT", 1.5, q=27), "full"),
    "ext_periodic_rlb_f32":     (lambda: ext_periodic(8, 7, 9, 30, "RLB", 1.5, prec="f32"), "full"),
    "ext_interleaved":          (lambda: ext_interleaved(), "full"),
    # dynamics-based boundary managers (createDynamicsBasedLocalBoundaryCondition3D / createEquilibriumBoundaryCondition3D): the plane
    # dynamics completes the populations at collision time, nothing runs after streaming
    "cavity_dynlocal_10x9x8":   (lambda: cavity(10, 9, 8, 40, model="dynlocal", every=20), "full"),
    "cavity_equilibrium_rlb":   (lambda: cavity(9, 8, 10, 30, dyn="RLB", model="equilibrium"), "full"),
    "cavity_dynlocal_q27_f32":  (lambda: cavity(8, 8, 9, 30, q=27, prec="f32", model="dynlocal"), "full"),
    "cavity_dynlocal_incbgk":   (lambda: cavity(8, 9, 10, 30, dyn="INCBGK", model="dynlocal"), "full"),
    "obstacle_dynlocal":        (lambda: obstacle(20, 10, 9, 40, omega=1.4, model="dynlocal"), "full"),
    "inlet_outlet_mixed_models": (lambda: inlet_outlet_mixed(12, 9, 8, 40), "full"),
    # createInterpBoundaryCondition3D: plain Dirichlet dynamics on the planes + the finite-difference plane processor
    "cavity_interp_10x9x8":     (lambda: cavity(10, 9, 8, 40, model="interp", every=20), "full"),
    "cavity_interp_rlb_q27":    (lambda: cavity(8, 8, 9, 30, dyn="RLB", q=27, model="interp"), "full"),
    "cavity_interp_incbgk_f32": (lambda: cavity(9, 8, 10, 30, dyn="INCBGK", prec="f32", model="interp"), "full"),
    "obstacle_interp":          (lambda: obstacle(20, 10, 9, 40, omega=1.4, model="interp"), "full"),
    # edge cases
    "mixed_omega_11x8x9":       (lambda: mixed_omega(11, 8, 9, 40), "full"),
    "inlet_outlet_local":       (lambda: inlet_outlet_local(12, 9, 8, 40), "full"),
    "thin_4x5x3":               (lambda: cavity(4, 5, 3, 10), "full"),
    # defineDynamics over boundary cells after the boundary set-up (the "eraser tool")
    "cavity_eraser_10x9x8":     (lambda: cavity_eraser(10, 9, 8, 40), "full"),
    "cavity_eraser_zouhe_q27":  (lambda: cavity_eraser(9, 8, 8, 30, model="zouhe", q=27), "full"),
    "cavity_eraser_interp_f32": (lambda: cavity_eraser(9, 8, 10, 30, model="interp", prec="f32"), "full"),
}

# Bench-scale cases: large enough that both halves of the sweep run several tiles per block in their
# grid-stride loop and the L2-prefetch branch sees a real stride (> 148 SMs x 128 blocks x 512 cells = 9.7 M cells).  Goldens
# (sha256 + probe cells) come from the unmodified reference like all others; the C oracle replays them in ~30 s each.
LARGE_CASES = {
    "cavity_bgk_256_10":              (lambda: cavity(256, 256, 256, 10, dumps=False).dump("f_end").dump_rho_u("rho_u_end"), "probe"),
    "channel_smago_q27_f32_224x160x320": (lambda: channel(224, 160, 320, 6, dyn="SMAGO", omega=1.99, q=27, cs=0.14, perturb=0.01,
                                                            prec="f32"), "probe"),
}

PROBE_CELLS = 97  # cells sampled (deterministically) from large dumps


def probe_index(shape, n=PROBE_CELLS):
    """Deterministic sample of cells: corners, a face/edge walk and an LCG scatter."""
    nx, ny, nz = shape[:3]
    total = nx * ny * nz
    idx = {0, total - 1, nz - 1, (ny - 1) * nz, (nx - 1) * ny * nz, (nx // 2 * ny + ny // 2) * nz + nz // 2}
    s = 12345
    while len(idx) < min(n, total):
        s = (1103515245 * s + 12345) % (1 << 31)
        idx.add(s % total)
    return np.array(sorted(idx), dtype=np.int64)
