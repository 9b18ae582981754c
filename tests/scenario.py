"""Scenario replay: one list of set-up / run operations, three engines.

A :class:`Scenario` records calls that read like the reference's own user code
(``MultiBlockLattice3D(nx,ny,nz,new BGKdynamics(omega))``, ``createLocalBoundaryCondition3D()
->setVelocityConditionOnBlockBoundaries``, ``setBoundaryVelocity``, ``initializeAtEquilibrium``,
``lattice.initialize()``, ``collideAndStream()`` -- cf. jlabos/src/testing/test_cylinder.py:100-119 of
the reference) and replays them on

* the UNMODIFIED reference, through ``oracle/_ref/ref_driver`` (only where /root/reference was
  available to build it; used by oracle/gen_golden.py to produce tests/golden/),
* the CPU oracle ``oracle/liboracle.so`` (:class:`OracleEngine`),
* the CUDA engine ``libplb_b200.so`` through its C-ABI (:class:`palabos_b200.Lattice`).

Test infrastructure only.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DRIVER = os.path.join(ORACLE_DIR, "_ref", "ref_driver")
REF_DRIVER_F32 = os.path.join(ORACLE_DIR, "_ref", "ref_driver_f32")

DYN = {"BGK": 0, "RLB": 1, "TRT": 2, "SMAGO": 3, "BB": 4, "NODYN": 5, "INCBGK": 6}
BC = {"local": 0, "zouhe": 1, "dynlocal": 2, "equilibrium": 3, "interp": 4}


def full_box(nx, ny, nz):
    return (0, nx - 1, 0, ny - 1, 0, nz - 1)


class Scenario:
    def __init__(self, q, prec, nx, ny, nz, dyn, p1, p2=0.0, blocks=(1, 1, 1)):
        assert q in (19, 27) and prec in ("f64", "f32")
        self.q, self.prec, self.nx, self.ny, self.nz = q, prec, nx, ny, nz
        self.dyn, self.p1, self.p2, self.blocks = dyn, float(p1), float(p2), tuple(blocks)
        self.ops = []

    @property
    def dtype(self):
        return np.float64 if self.prec == "f64" else np.float32

    @property
    def box(self):
        return full_box(self.nx, self.ny, self.nz)

    def _add(self, *op):
        self.ops.append(op)
        return self

    def periodic(self, px, py, pz):
        return self._add("periodic", int(px), int(py), int(pz))

    def dynamics(self, box, kind, p1, p2=0.0):
        return self._add("dynamics", tuple(box), kind, float(p1), float(p2))

    def dynamics_mask(self, mask, flag, kind, p1, p2=0.0):
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        assert mask.shape == (self.nx, self.ny, self.nz)
        return self._add("dynamics_mask", mask, int(flag), kind, float(p1), float(p2))

    def bc_box(self, model, block=None, app=None):
        block = tuple(block or self.box)
        return self._add("bc_box", model, block, tuple(app or block))

    def bc_vel(self, model, direction, orientation, box):
        return self._add("bc_vel", model, direction, orientation, tuple(box))

    def bc_pres(self, model, direction, orientation, box):
        return self._add("bc_pres", model, direction, orientation, tuple(box))

    def bc_edge(self, model, plane, n1, n2, box):
        return self._add("bc_edge", model, plane, n1, n2, tuple(box))

    def bc_corner(self, model, xn, yn, zn, x, y, z):
        return self._add("bc_corner", model, xn, yn, zn, x, y, z)

    def bc_iedge(self, model, plane, n1, n2, box):
        """addInternalVelocityEdge<plane><n1><n2>: a concave edge"""
        return self._add("bc_iedge", model, plane, n1, n2, tuple(box))

    def bc_icorner(self, model, xn, yn, zn, x, y, z):
        return self._add("bc_icorner", model, xn, yn, zn, x, y, z)

    def set_velocity(self, box, u):
        return self._add("set_velocity", tuple(box), tuple(float(v) for v in u))

    def set_density(self, box, rho):
        return self._add("set_density", tuple(box), float(rho))

    def init_eq(self, box, rho, u):
        return self._add("init_eq", tuple(box), float(rho), tuple(float(v) for v in u))

    def init_eq_field(self, box, rho_u):
        rho_u = np.ascontiguousarray(rho_u, dtype=np.float64)
        return self._add("init_eq_field", tuple(box), rho_u)

    def upload(self, pops):
        pops = np.ascontiguousarray(pops, dtype=self.dtype)
        assert pops.shape == (self.nx, self.ny, self.nz, self.q)
        return self._add("upload", pops)

    def initialize(self):
        return self._add("initialize")

    def run(self, n):
        return self._add("run", int(n))

    def stats(self):
        return self._add("stats")

    def dump(self, name):
        return self._add("dump", name)

    def dump_rho_u(self, name):
        return self._add("dump_rho_u", name)

    # ---- the external-moment sweep of the fork's application (src/variables.hh:475-486, viscosityTest/src/test.cpp:64-65)

    @property
    def external(self):
        return any(op[0] == "ext_integrate" for op in self.ops)

    @property
    def periodicity(self):
        """Periodicity in force when the fields are integrated (the enveloped engines fix it at creation)."""
        per = (0, 0, 0)
        for op in self.ops:
            if op[0] == "periodic":
                per = op[1:]
            if op[0] == "ext_integrate":
                break
        return per

    def ext_integrate(self):
        """rhoBar / j fields + ExternalRhoJcollideAndStream3D at level 0 + BoxRhoBarJfunctional3D at level 3; must precede
        the boundary set-up (whose level-0 processors then run after the sweep) and follow the periodicity toggles."""
        assert all(op[0] == "periodic" for op in self.ops), "ext_integrate comes right after the periodicity toggles"
        return self._add("ext_integrate")

    def ext_apply(self, box=None):
        return self._add("ext_apply", tuple(box or self.box))

    def ext_exec(self, n):
        return self._add("ext_exec", int(n))

    def ext_add_j(self, box, dj):
        return self._add("ext_add_j", tuple(box), tuple(float(v) for v in dj))

    def dump_rhobar_j(self, name):
        return self._add("dump_rhobar_j", name)


# ----------------------------------------------------------------------------- reference


def have_reference_driver(prec="f64"):
    return os.path.exists(REF_DRIVER if prec == "f64" else REF_DRIVER_F32)


def run_reference(sc: Scenario, timed_steps=None):
    """Replay on the real reference (subprocess).  Returns {'stats': [k,6] array, name: array}."""
    fmt = lambda t: " ".join(repr(v) if isinstance(v, float) else str(v) for v in t)
    with tempfile.TemporaryDirectory() as tmp:
        lines = ["blocks %d %d %d" % sc.blocks,
                 "lattice D3Q%d %s %d %d %d %s %r %r" % (sc.q, sc.prec, sc.nx, sc.ny, sc.nz, sc.dyn, sc.p1, sc.p2)]
        blob = 0
        dumps = []
        files = []
        for op in sc.ops:
            k = op[0]
            if k == "periodic":
                lines.append("periodic %d %d %d" % op[1:])
            elif k == "dynamics":
                lines.append("dynamics %s %s %r %r" % (fmt(op[1]), op[2], op[3], op[4]))
            elif k == "dynamics_mask":
                fn = "blob%d.bin" % blob; blob += 1
                op[1].tofile(os.path.join(tmp, fn))
                lines.append("dynamics_mask %s %d %s %r %r" % (fn, op[2], op[3], op[4], op[5]))
            elif k == "bc_box":
                lines.append("bc_box %s %s %s" % (op[1], fmt(op[2]), fmt(op[3])))
            elif k in ("bc_vel", "bc_pres"):
                lines.append("%s %s %d %d %s" % (k, op[1], op[2], op[3], fmt(op[4])))
            elif k == "bc_edge":
                lines.append("bc_edge %s %d %d %d %s" % (op[1], op[2], op[3], op[4], fmt(op[5])))
            elif k == "bc_corner":
                lines.append("bc_corner %s %s" % (op[1], fmt(op[2:])))
            elif k == "bc_iedge":
                lines.append("bc_iedge %s %d %d %d %s" % (op[1], op[2], op[3], op[4], fmt(op[5])))
            elif k == "bc_icorner":
                lines.append("bc_icorner %s %s" % (op[1], fmt(op[2:])))
            elif k == "set_velocity":
                lines.append("set_velocity %s %s" % (fmt(op[1]), fmt(op[2])))
            elif k == "set_density":
                lines.append("set_density %s %r" % (fmt(op[1]), op[2]))
            elif k == "init_eq":
                lines.append("init_eq %s %r %s" % (fmt(op[1]), op[2], fmt(op[3])))
            elif k == "init_eq_field":
                fn = "blob%d.bin" % blob; blob += 1
                op[2].tofile(os.path.join(tmp, fn))
                lines.append("init_eq_field %s %s" % (fmt(op[1]), fn))
            elif k == "upload":
                fn = "blob%d.bin" % blob; blob += 1
                op[1].tofile(os.path.join(tmp, fn))
                lines.append("upload %s" % fn)
            elif k == "initialize":
                lines.append("initialize")
            elif k == "ext_integrate":
                lines.append("ext_integrate")
            elif k == "ext_apply":
                lines.append("ext_apply %s" % fmt(op[1]))
            elif k == "ext_exec":
                lines.append("ext_exec %d" % op[1])
            elif k == "ext_add_j":
                lines.append("ext_add_j %s %s" % (fmt(op[1]), fmt(op[2])))
            elif k == "dump_rhobar_j":
                dumps.append((k, op[1]))
                lines.append("dump_rhobar_j %s.bin" % op[1])
            elif k == "run":
                lines.append("run %d" % op[1])
            elif k == "stats":
                lines.append("stats")
            elif k in ("dump", "dump_rho_u"):
                dumps.append((k, op[1]))
                lines.append("%s %s.bin" % (k, op[1]))
            elif k == "ref_file":
                # a command of the reference driver that writes a file (vtk, save_block): op = (k, command line, file name, key)
                lines.append(op[1])
                files.append((op[2], op[3]))
            else:
                raise ValueError(k)
        if timed_steps:
            lines.append("time_run %d" % timed_steps)
        with open(os.path.join(tmp, "scenario.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
        exe = REF_DRIVER if sc.prec == "f64" else REF_DRIVER_F32
        out = subprocess.run([exe, "scenario.txt"], cwd=tmp, check=True, capture_output=True, text=True).stdout
        res = {"stats": [], "time": None}
        for ln in out.splitlines():
            w = ln.split()
            if w and w[0] == "STATS":
                res["stats"].append([float(v) for v in w[1:]])
            if w and w[0] == "TIME":
                res["time"] = (int(w[1]), float(w[2]), float(w[3]))
        res["stats"] = np.array(res["stats"], dtype=np.float64).reshape(-1, 6)
        for kind, name in dumps:
            a = np.fromfile(os.path.join(tmp, name + ".bin"), dtype=sc.dtype)
            res[name] = a.reshape(sc.nx, sc.ny, sc.nz, sc.q if kind == "dump" else 4)   # dump_rho_u, dump_rhobar_j: 4
        for fn, key in files:
            res[key] = np.frombuffer(open(os.path.join(tmp, fn), "rb").read(), dtype=np.uint8)
        return res


# ----------------------------------------------------------------------------- oracle


_oracle_lib = None


def oracle_lib():
    """ctypes handle on oracle/liboracle.so (built on demand: `make -C oracle liboracle`)."""
    global _oracle_lib
    if _oracle_lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("plb_oracle.c", "plb_oracle_impl.h", "plb_oracle.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle"], check=True, capture_output=True)
        lib = ctypes.CDLL(so)
        lib.plbo_create.restype = ctypes.c_void_p
        lib.plbo_create.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double] * 2
        lib.plbo_last_error.restype = ctypes.c_char_p
        _oracle_lib = lib
    return _oracle_lib


def _box(b):
    return (ctypes.c_int * 6)(*[int(v) for v in b])


def _vec(u):
    return (ctypes.c_double * 3)(*[float(v) for v in u])


class OracleEngine:
    """The CPU oracle behind the same method names as palabos_b200.Lattice."""

    def __init__(self, q, prec, nx, ny, nz, dyn, p1, p2=0.0):
        self.lib = oracle_lib()
        self.q, self.prec, self.nx, self.ny, self.nz = q, prec, nx, ny, nz
        self.dtype = np.float64 if prec == "f64" else np.float32
        self.h = ctypes.c_void_p(self.lib.plbo_create(q, 8 if prec == "f64" else 4, nx, ny, nz, DYN[dyn], p1, p2))
        if not self.h:
            raise RuntimeError(self.lib.plbo_last_error().decode())

    def close(self):
        if self.h:
            self.lib.plbo_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, ok):
        if not ok:
            raise RuntimeError(self.lib.plbo_last_error().decode())

    def set_periodicity(self, px, py, pz):
        self._ck(self.lib.plbo_set_periodicity(self.h, int(px), int(py), int(pz)))

    def set_statistics_box(self, box):
        self._ck(self.lib.plbo_set_statistics_box(self.h, _box(box)))

    def define_dynamics(self, box, kind, p1, p2=0.0):
        self._ck(self.lib.plbo_define_dynamics_box(self.h, _box(box), DYN[kind], ctypes.c_double(p1), ctypes.c_double(p2)))

    def define_dynamics_mask(self, mask, flag, kind, p1, p2=0.0):
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        self._ck(self.lib.plbo_define_dynamics_mask(self.h, mask.ctypes.data_as(ctypes.c_void_p), int(flag), DYN[kind],
                                                    ctypes.c_double(p1), ctypes.c_double(p2)))

    def set_velocity_condition_on_block_boundaries(self, model, block, app):
        self._ck(self.lib.plbo_set_velocity_condition_on_block_boundaries(self.h, BC[model], _box(block), _box(app)))

    def add_velocity_boundary(self, model, d, o, box):
        self._ck(self.lib.plbo_add_velocity_boundary(self.h, BC[model], d, o, _box(box)))

    def add_pressure_boundary(self, model, d, o, box):
        self._ck(self.lib.plbo_add_pressure_boundary(self.h, BC[model], d, o, _box(box)))

    def add_external_velocity_edge(self, model, plane, n1, n2, box):
        self._ck(self.lib.plbo_add_external_velocity_edge(self.h, BC[model], plane, n1, n2, _box(box)))

    def add_external_velocity_corner(self, model, xn, yn, zn, x, y, z):
        self._ck(self.lib.plbo_add_external_velocity_corner(self.h, BC[model], xn, yn, zn, x, y, z))

    def add_internal_velocity_edge(self, model, plane, n1, n2, box):
        self._ck(self.lib.plbo_add_internal_velocity_edge(self.h, BC[model], plane, n1, n2, _box(box)))

    def add_internal_velocity_corner(self, model, xn, yn, zn, x, y, z):
        self._ck(self.lib.plbo_add_internal_velocity_corner(self.h, BC[model], xn, yn, zn, x, y, z))

    def set_boundary_velocity(self, box, u):
        self._ck(self.lib.plbo_set_boundary_velocity(self.h, _box(box), _vec(u)))

    def set_boundary_density(self, box, rho):
        self._ck(self.lib.plbo_set_boundary_density(self.h, _box(box), ctypes.c_double(rho)))

    def initialize_at_equilibrium(self, box, rho, u):
        self._ck(self.lib.plbo_initialize_at_equilibrium(self.h, _box(box), ctypes.c_double(rho), _vec(u)))

    def initialize_at_equilibrium_field(self, box, rho_u):
        rho_u = np.ascontiguousarray(rho_u, dtype=np.float64)
        self._ck(self.lib.plbo_initialize_at_equilibrium_field(self.h, _box(box), rho_u.ctypes.data_as(ctypes.c_void_p)))

    def upload(self, pops, box=None):
        box = box or full_box(self.nx, self.ny, self.nz)
        pops = np.ascontiguousarray(pops, dtype=self.dtype)
        self._ck(self.lib.plbo_upload(self.h, _box(box), pops.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(pops.nbytes)))

    def download(self, box=None):
        box = box or full_box(self.nx, self.ny, self.nz)
        shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, self.q)
        out = np.empty(shape, dtype=self.dtype)
        self._ck(self.lib.plbo_download(self.h, _box(box), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.nbytes)))
        return out

    def initialize(self):
        self._ck(self.lib.plbo_initialize(self.h))

    def collide_and_stream(self, n=1):
        self._ck(self.lib.plbo_collide_and_stream(self.h, int(n)))

    def enable_external_rhoj(self):
        self._ck(self.lib.plbo_enable_external_rhoj(self.h))

    def execute_internal_processors(self, n=1):
        self._ck(self.lib.plbo_execute_internal_processors(self.h, int(n)))

    def apply_rhobar_j(self, box=None):
        self._ck(self.lib.plbo_apply_rhobar_j(self.h, _box(box or full_box(self.nx, self.ny, self.nz))))

    def upload_rhobar_j(self, fld, box=None):
        box = box or full_box(self.nx, self.ny, self.nz)
        fld = np.ascontiguousarray(fld, dtype=self.dtype)
        self._ck(self.lib.plbo_upload_rhobar_j(self.h, _box(box), fld.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(fld.nbytes)))

    def download_rhobar_j(self, box=None):
        box = box or full_box(self.nx, self.ny, self.nz)
        out = np.empty((box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, 4), dtype=self.dtype)
        self._ck(self.lib.plbo_download_rhobar_j(self.h, _box(box), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.nbytes)))
        return out

    def stored_statistics(self):
        out = (ctypes.c_double * 3)()
        n = ctypes.c_uint64()
        self._ck(self.lib.plbo_get_stored_statistics(self.h, out, ctypes.byref(n)))
        return out[0], out[1], out[2], n.value

    def compute_average_energy(self, box=None):
        out = ctypes.c_double()
        self._ck(self.lib.plbo_compute_average_energy(self.h, _box(box or full_box(self.nx, self.ny, self.nz)), ctypes.byref(out)))
        return out.value

    def compute_average_density(self, box=None):
        out = ctypes.c_double()
        self._ck(self.lib.plbo_compute_average_density(self.h, _box(box or full_box(self.nx, self.ny, self.nz)), ctypes.byref(out)))
        return out.value

    def compute_density_velocity(self, box=None):
        box = box or full_box(self.nx, self.ny, self.nz)
        shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, 4)
        out = np.empty(shape, dtype=self.dtype)
        self._ck(self.lib.plbo_compute_density_velocity(self.h, _box(box), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.nbytes)))
        return out


# ----------------------------------------------------------------------------- replay


def stats_row(engine, it, prec):
    """Same six numbers as ref_driver's STATS line: iter, stored avg density, stored avg energy,
    stored max velocity (multiBlockLattice3D.hh:531-548), computeAverageEnergy, computeAverageDensity."""
    av_rhobar, av_usqr, max_usqr, _ = engine.stored_statistics()
    if prec == "f32":   # Descriptor<float>::fullRho takes and returns float (multiBlockLattice3D.hh:532-535)
        rho = float(np.float32(np.float32(av_rhobar) + np.float32(1)))
    else:
        rho = av_rhobar + 1.0
    with np.errstate(invalid="ignore"):
        maxu = float(np.sqrt(np.float64(max_usqr)))
    return [it, rho, 0.5 * av_usqr, maxu, engine.compute_average_energy(), engine.compute_average_density()]


def replay(sc: Scenario, engine):
    """Replay `sc` on an engine object (OracleEngine or palabos_b200.Lattice)."""
    res = {"stats": []}
    it = 0
    for op in sc.ops:
        k = op[0]
        if k == "periodic":
            engine.set_periodicity(*op[1:])
        elif k == "dynamics":
            engine.define_dynamics(op[1], op[2], op[3], op[4])
        elif k == "dynamics_mask":
            engine.define_dynamics_mask(op[1], op[2], op[3], op[4], op[5])
        elif k == "bc_box":
            engine.set_velocity_condition_on_block_boundaries(op[1], op[2], op[3])
        elif k == "bc_vel":
            engine.add_velocity_boundary(op[1], op[2], op[3], op[4])
        elif k == "bc_pres":
            engine.add_pressure_boundary(op[1], op[2], op[3], op[4])
        elif k == "bc_edge":
            engine.add_external_velocity_edge(op[1], op[2], op[3], op[4], op[5])
        elif k == "bc_corner":
            engine.add_external_velocity_corner(*op[1:])
        elif k == "bc_iedge":
            engine.add_internal_velocity_edge(op[1], op[2], op[3], op[4], op[5])
        elif k == "bc_icorner":
            engine.add_internal_velocity_corner(*op[1:])
        elif k == "set_velocity":
            engine.set_boundary_velocity(op[1], op[2])
        elif k == "set_density":
            engine.set_boundary_density(op[1], op[2])
        elif k == "init_eq":
            engine.initialize_at_equilibrium(op[1], op[2], op[3])
        elif k == "init_eq_field":
            engine.initialize_at_equilibrium_field(op[1], op[2])
        elif k == "upload":
            engine.upload(op[1])
        elif k == "initialize":
            engine.initialize()
        elif k == "ext_integrate":
            engine.enable_external_rhoj()
        elif k == "ext_apply":
            engine.apply_rhobar_j(op[1])
        elif k == "ext_exec":
            engine.execute_internal_processors(op[1])
            it += op[1]
        elif k == "ext_add_j":
            fld = engine.download_rhobar_j(op[1])
            fld[..., 1:] += np.asarray(op[2], dtype=fld.dtype)
            engine.upload_rhobar_j(fld, op[1])
        elif k == "dump_rhobar_j":
            res[op[1]] = engine.download_rhobar_j()
        elif k == "run":
            engine.collide_and_stream(op[1])
            it += op[1]
        elif k == "stats":
            res["stats"].append(stats_row(engine, it, sc.prec))
        elif k == "dump":
            res[op[1]] = engine.download()
        elif k == "dump_rho_u":
            res[op[1]] = engine.compute_density_velocity()
        elif k == "ref_file":
            pass                                  # only the reference driver writes these
        else:
            raise ValueError(k)
    res["stats"] = np.array(res["stats"], dtype=np.float64).reshape(-1, 6)
    return res


class EnvelopedOracle:
    """The oracle lattice of a scenario that uses the external-moment sweep: one cell larger on every non-periodic
    side (the reference's envelope, which ExternalRhoJcollideAndStream3D collides and streams,
    basicDynamics/dynamicsProcessor3D.hh:118-119), user coordinates shifted accordingly.  The CUDA engine does the same
    shift behind plb_b200_create_enveloped."""

    def __init__(self, q, prec, nx, ny, nz, dyn, p1, p2, periodic):
        self.per = tuple(int(bool(p)) for p in periodic)
        self.off = tuple(0 if p else 1 for p in self.per)
        self.nx, self.ny, self.nz, self.q = nx, ny, nz, q
        self.inner = OracleEngine(q, prec, nx + 2 * self.off[0], ny + 2 * self.off[1], nz + 2 * self.off[2], dyn, p1, p2)
        self.inner.set_periodicity(*self.per)
        self.inner.set_statistics_box(self._b(full_box(nx, ny, nz)))
        self.dtype = self.inner.dtype

    def close(self):
        self.inner.close()

    def _b(self, box):
        o = self.off
        box = box or full_box(self.nx, self.ny, self.nz)
        return (box[0] + o[0], box[1] + o[0], box[2] + o[1], box[3] + o[1], box[4] + o[2], box[5] + o[2])

    def set_periodicity(self, px, py, pz):
        if (int(bool(px)), int(bool(py)), int(bool(pz))) != self.per:
            raise RuntimeError("an enveloped lattice keeps the periodicity it was created with")

    def define_dynamics(self, box, *a):
        self.inner.define_dynamics(self._b(box), *a)

    def define_dynamics_mask(self, mask, flag, *a):
        o = self.off
        big = np.full((self.inner.nx, self.inner.ny, self.inner.nz), flag + 1, dtype=np.int32)
        big[o[0]:o[0] + self.nx, o[1]:o[1] + self.ny, o[2]:o[2] + self.nz] = mask
        self.inner.define_dynamics_mask(big, flag, *a)

    def set_velocity_condition_on_block_boundaries(self, model, block, app):
        self.inner.set_velocity_condition_on_block_boundaries(model, self._b(block), self._b(app))

    def add_velocity_boundary(self, model, d, o, box):
        self.inner.add_velocity_boundary(model, d, o, self._b(box))

    def add_pressure_boundary(self, model, d, o, box):
        self.inner.add_pressure_boundary(model, d, o, self._b(box))

    def add_external_velocity_edge(self, model, plane, n1, n2, box):
        self.inner.add_external_velocity_edge(model, plane, n1, n2, self._b(box))

    def add_external_velocity_corner(self, model, xn, yn, zn, x, y, z):
        self.inner.add_external_velocity_corner(model, xn, yn, zn, x + self.off[0], y + self.off[1], z + self.off[2])

    def add_internal_velocity_edge(self, model, plane, n1, n2, box):
        self.inner.add_internal_velocity_edge(model, plane, n1, n2, self._b(box))

    def add_internal_velocity_corner(self, model, xn, yn, zn, x, y, z):
        self.inner.add_internal_velocity_corner(model, xn, yn, zn, x + self.off[0], y + self.off[1], z + self.off[2])

    def set_boundary_velocity(self, box, u):
        self.inner.set_boundary_velocity(self._b(box), u)

    def set_boundary_density(self, box, rho):
        self.inner.set_boundary_density(self._b(box), rho)

    def initialize_at_equilibrium(self, box, rho, u):
        self.inner.initialize_at_equilibrium(self._b(box), rho, u)

    def initialize_at_equilibrium_field(self, box, rho_u):
        self.inner.initialize_at_equilibrium_field(self._b(box), rho_u)

    def upload(self, pops, box=None):
        self.inner.upload(pops, self._b(box))

    def download(self, box=None):
        return self.inner.download(self._b(box))

    def initialize(self):
        # MultiBlock3D::initialize = executeInternalProcessors (multiBlock/multiBlock3D.cpp:330-332): with the sweep
        # integrated as a processor this is a whole step
        self.inner.execute_internal_processors(1)

    def collide_and_stream(self, n=1):
        raise RuntimeError("a lattice with the external-moment sweep is stepped by execute_internal_processors")

    def enable_external_rhoj(self):
        self.inner.enable_external_rhoj()

    def execute_internal_processors(self, n=1):
        self.inner.execute_internal_processors(n)

    def apply_rhobar_j(self, box=None):
        self.inner.apply_rhobar_j(self._b(box))

    def upload_rhobar_j(self, fld, box=None):
        self.inner.upload_rhobar_j(fld, self._b(box))

    def download_rhobar_j(self, box=None):
        return self.inner.download_rhobar_j(self._b(box))

    def stored_statistics(self):
        return self.inner.stored_statistics()

    def compute_average_energy(self, box=None):
        return self.inner.compute_average_energy(self._b(box))

    def compute_average_density(self, box=None):
        return self.inner.compute_average_density(self._b(box))

    def compute_density_velocity(self, box=None):
        return self.inner.compute_density_velocity(self._b(box))


def run_oracle(sc: Scenario):
    if sc.external:
        eng = EnvelopedOracle(sc.q, sc.prec, sc.nx, sc.ny, sc.nz, sc.dyn, sc.p1, sc.p2, sc.periodicity)
    else:
        eng = OracleEngine(sc.q, sc.prec, sc.nx, sc.ny, sc.nz, sc.dyn, sc.p1, sc.p2)
    try:
        return replay(sc, eng)
    finally:
        eng.close()


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def max_rel_diff(a, b):
    """max |a-b| / max|b| -- the 'relative' of north_star's 1e-10 (population scale)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))
