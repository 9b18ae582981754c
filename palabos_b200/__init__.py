"""palabos_b200 -- ctypes binding of libplb_b200.so (include/plb_b200.h).

The product is the C-ABI shared library (hand-written sm_100a CUDA behind the Palabos API for
MultiBlockLattice3D::collideAndStream and its boundary processors / envelope exchange /
statistics); the C++ mirror of the reference's classes is include/palabos_b200/.  This module is
the thin Python view of the same ABI used by tests/ and bench.py: :class:`Lattice` has one method
per C entry point, named after the reference wrapper it replaces.

There is no CPU fallback: creating a lattice without the built library or without a CUDA device
raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(ROOT, "libplb_b200.so")
CSRC = os.path.join(ROOT, "csrc")

DYN = {"BGK": 0, "RLB": 1, "TRT": 2, "SMAGO": 3, "BB": 4, "NODYN": 5, "INCBGK": 6}
BC = {"local": 0, "zouhe": 1, "dynlocal": 2, "equilibrium": 3, "interp": 4}

#: every symbol include/plb_b200.h declares (tests check the library exports them all)
SYMBOLS = [
    "plb_b200_create", "plb_b200_create_distributed", "plb_b200_destroy", "plb_b200_last_error",
    "plb_b200_abi_version", "plb_b200_comm_unique_id", "plb_b200_comm_init", "plb_b200_slab_range",
    "plb_b200_set_stream", "plb_b200_set_periodicity", "plb_b200_toggle_internal_statistics",
    "plb_b200_define_dynamics_box", "plb_b200_define_dynamics_mask", "plb_b200_add_velocity_boundary",
    "plb_b200_add_pressure_boundary", "plb_b200_add_external_velocity_edge",
    "plb_b200_add_external_velocity_corner", "plb_b200_add_internal_velocity_edge", "plb_b200_add_internal_velocity_corner",
    "plb_b200_set_velocity_condition_on_block_boundaries",
    "plb_b200_set_boundary_velocity", "plb_b200_set_boundary_density", "plb_b200_initialize_at_equilibrium",
    "plb_b200_initialize_at_equilibrium_field", "plb_b200_initialize", "plb_b200_upload", "plb_b200_download",
    "plb_b200_local_bytes", "plb_b200_collide_and_stream", "plb_b200_synchronize",
    "plb_b200_get_stored_statistics", "plb_b200_compute_average_energy", "plb_b200_compute_average_density",
    "plb_b200_compute_density_velocity", "plb_b200_launch_count", "plb_b200_device_bytes",
    "plb_b200_boundary_stages", "plb_b200_profile_sweeps", "plb_b200_profile_read", "plb_b200_halo_transport", "plb_b200_coproc_add_domain", "plb_b200_coproc_send",
    "plb_b200_coproc_receive", "plb_b200_coproc_collide_and_stream", "plb_b200_coproc_remove_domain",
    "plb_b200_save_binary_block", "plb_b200_load_binary_block", "plb_b200_parallel_save", "plb_b200_parallel_load",
    "plb_b200_create_enveloped", "plb_b200_execute_internal_processors", "plb_b200_apply_rhobar_j",
    "plb_b200_upload_rhobar_j", "plb_b200_download_rhobar_j",
    "plb_b200_voxelize", "plb_b200_stl_read", "plb_b200_stl_free", "plb_b200_mesh_to_lattice",
]


def build(verbose=False):
    """Compile libplb_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    jobs = str(min(8, os.cpu_count() or 1))
    r = subprocess.run(["make", "-C", CSRC, "-j", jobs], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libplb_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def lib():
    """The loaded library.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("palabos_b200/libplb_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.plb_b200_create.restype = vp
        L.plb_b200_create.argtypes = [ci] * 6 + [cd] * 2
        L.plb_b200_create_distributed.restype = vp
        L.plb_b200_create_distributed.argtypes = [ci] * 6 + [cd] * 2 + [ci] * 3
        L.plb_b200_create_enveloped.restype = vp
        L.plb_b200_create_enveloped.argtypes = [ci] * 6 + [cd] * 2 + [ci] * 6
        L.plb_b200_execute_internal_processors.argtypes = [vp, ci]
        L.plb_b200_apply_rhobar_j.argtypes = [vp, vp]
        L.plb_b200_upload_rhobar_j.argtypes = [vp, vp, vp, ctypes.c_size_t]
        L.plb_b200_download_rhobar_j.argtypes = [vp, vp, vp, ctypes.c_size_t]
        L.plb_b200_destroy.restype = None
        L.plb_b200_destroy.argtypes = [vp]
        L.plb_b200_last_error.restype = ctypes.c_char_p
        L.plb_b200_local_bytes.restype = ctypes.c_size_t
        L.plb_b200_local_bytes.argtypes = [vp, vp]
        L.plb_b200_launch_count.restype = ctypes.c_uint64
        L.plb_b200_launch_count.argtypes = [vp]
        L.plb_b200_device_bytes.restype = ctypes.c_size_t
        L.plb_b200_device_bytes.argtypes = [vp]
        L.plb_b200_comm_unique_id.argtypes = [vp]
        L.plb_b200_comm_init.argtypes = [vp, vp]
        L.plb_b200_set_stream.argtypes = [vp, vp]
        L.plb_b200_set_periodicity.argtypes = [vp, ci, ci, ci]
        L.plb_b200_toggle_internal_statistics.argtypes = [vp, ci]
        L.plb_b200_define_dynamics_box.argtypes = [vp, vp, ci, cd, cd]
        L.plb_b200_define_dynamics_mask.argtypes = [vp, vp, ci, ci, cd, cd]
        L.plb_b200_add_velocity_boundary.argtypes = [vp, ci, ci, ci, vp]
        L.plb_b200_add_pressure_boundary.argtypes = [vp, ci, ci, ci, vp]
        L.plb_b200_add_external_velocity_edge.argtypes = [vp, ci, ci, ci, ci, vp]
        L.plb_b200_add_external_velocity_corner.argtypes = [vp] + [ci] * 7
        L.plb_b200_add_internal_velocity_edge.argtypes = [vp, ci, ci, ci, ci, vp]
        L.plb_b200_add_internal_velocity_corner.argtypes = [vp] + [ci] * 7
        L.plb_b200_set_velocity_condition_on_block_boundaries.argtypes = [vp, ci, vp, vp]
        L.plb_b200_set_boundary_velocity.argtypes = [vp, vp, vp]
        L.plb_b200_set_boundary_density.argtypes = [vp, vp, cd]
        L.plb_b200_initialize_at_equilibrium.argtypes = [vp, vp, cd, vp]
        L.plb_b200_initialize_at_equilibrium_field.argtypes = [vp, vp, vp]
        L.plb_b200_initialize.argtypes = [vp]
        L.plb_b200_upload.argtypes = [vp, vp, vp, ctypes.c_size_t]
        L.plb_b200_download.argtypes = [vp, vp, vp, ctypes.c_size_t]
        L.plb_b200_collide_and_stream.argtypes = [vp, ci]
        L.plb_b200_synchronize.argtypes = [vp]
        L.plb_b200_get_stored_statistics.argtypes = [vp, vp, vp]
        L.plb_b200_compute_average_energy.argtypes = [vp, vp, vp]
        L.plb_b200_compute_average_density.argtypes = [vp, vp, vp]
        L.plb_b200_compute_density_velocity.argtypes = [vp, vp, vp, ctypes.c_size_t]
        L.plb_b200_boundary_stages.argtypes = [vp, vp, vp]
        L.plb_b200_profile_sweeps.argtypes = [vp, ci]
        L.plb_b200_profile_read.argtypes = [vp, vp, vp]
        L.plb_b200_halo_transport.restype = ctypes.c_char_p
        L.plb_b200_halo_transport.argtypes = [vp]
        L.plb_b200_slab_range.argtypes = [ci, ci, ci, vp, vp]
        L.plb_b200_coproc_add_domain.argtypes = [ci, ci, ci, ci, cd, vp]
        L.plb_b200_coproc_send.argtypes = [ci, vp, vp, ctypes.c_size_t]
        L.plb_b200_coproc_receive.argtypes = [ci, vp, vp, ctypes.c_size_t]
        L.plb_b200_coproc_collide_and_stream.argtypes = [ci]
        L.plb_b200_coproc_remove_domain.argtypes = [ci]
        L.plb_b200_save_binary_block.argtypes = [vp, ctypes.c_char_p]
        L.plb_b200_load_binary_block.argtypes = [vp, ctypes.c_char_p]
        L.plb_b200_parallel_save.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.plb_b200_parallel_load.argtypes = [vp, ctypes.c_char_p]
        L.plb_b200_voxelize.argtypes = [vp, ctypes.c_int64, ci, ci, ci, ci, vp, vp]
        L.plb_b200_stl_read.argtypes = [ctypes.c_char_p, vp, vp]
        L.plb_b200_stl_free.restype = None
        L.plb_b200_stl_free.argtypes = [vp]
        L.plb_b200_mesh_to_lattice.argtypes = [vp, ctypes.c_int64, ci, ci, ci, ci, cd, vp, vp, vp, vp]
        _lib = L
    return _lib


class PlbError(RuntimeError):
    """A C-ABI call returned 0 (the shim's equivalent of the reference's PlbLogicException)."""


def _box(b):
    return (ctypes.c_int * 6)(*[int(v) for v in b])


def _vec(u):
    return (ctypes.c_double * 3)(*[float(v) for v in u])


def _ck(ok):
    if not ok:
        raise PlbError(lib().plb_b200_last_error().decode())


def stl_read(path):
    """Triangles of an ASCII or binary STL file as float64 [n,3,3] (TriangleSet<T>(fileName, FLT))."""
    ptr, n = ctypes.c_void_p(), ctypes.c_int64()
    _ck(lib().plb_b200_stl_read(str(path).encode(), ctypes.byref(ptr), ctypes.byref(n)))
    try:
        buf = (ctypes.c_double * (n.value * 9)).from_address(ptr.value) if n.value else []
        return np.array(buf, dtype=np.float64).reshape(-1, 3, 3)
    finally:
        lib().plb_b200_stl_free(ptr)


def mesh_to_lattice(soup, resolution, reference_direction, margin, extra_layer=0, inflate=1e-3):
    """DEFscaledMesh + inflate: returns (triangles float64 [n,3,3] in lattice units, dims, dx, location)."""
    soup = np.ascontiguousarray(soup, dtype=np.float64).reshape(-1, 3, 3)
    tri = np.empty(soup.shape, dtype=np.float64)
    dims, dx, loc = (ctypes.c_int * 3)(), ctypes.c_double(), (ctypes.c_double * 3)()
    _ck(lib().plb_b200_mesh_to_lattice(soup.ctypes.data_as(ctypes.c_void_p), soup.shape[0], int(resolution), int(reference_direction),
                                       int(margin), int(extra_layer), float(inflate), tri.ctypes.data_as(ctypes.c_void_p),
                                       dims, ctypes.byref(dx), loc))
    return tri, tuple(dims), dx.value, tuple(loc)


def voxelize(triangles, dims, border_width=1, return_ms=False):
    """voxelize(mesh, domain, borderWidth) + detectBorderLine on the device: int32 [nx,ny,nz] of voxelFlag values."""
    tri = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 3, 3)
    nx, ny, nz = [int(d) for d in dims]
    flags = np.empty((nx, ny, nz), dtype=np.int32)
    ms = ctypes.c_double()
    _ck(lib().plb_b200_voxelize(tri.ctypes.data_as(ctypes.c_void_p), tri.shape[0], nx, ny, nz, int(border_width),
                                flags.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ms)))
    return (flags, ms.value) if return_ms else flags


def slab_range(nx, n_slabs, slab):
    x0, x1 = ctypes.c_int(), ctypes.c_int()
    if not lib().plb_b200_slab_range(nx, n_slabs, slab, ctypes.byref(x0), ctypes.byref(x1)):
        raise PlbError(lib().plb_b200_last_error().decode())
    return x0.value, x1.value


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    if not lib().plb_b200_comm_unique_id(buf):
        raise PlbError(lib().plb_b200_last_error().decode())
    return buf.raw


class Lattice:
    """One MultiBlockLattice3D<T,Descriptor> on the GPU(s), through the C-ABI."""

    def __init__(self, q, prec, nx, ny, nz, dyn, p1, p2=0.0, rank=0, n_ranks=1, slabs_per_rank=1, enveloped_periodic=None):
        """enveloped_periodic=(px,py,pz): a lattice stepped by the external-moment sweep (plb_b200_create_enveloped)"""
        self.lib = lib()
        self.q, self.prec, self.nx, self.ny, self.nz = q, prec, nx, ny, nz
        self.rank, self.n_ranks, self.slabs_per_rank = rank, n_ranks, slabs_per_rank
        self.dtype = np.float64 if prec == "f64" else np.float32
        if enveloped_periodic is not None:
            px, py, pz = [int(bool(p)) for p in enveloped_periodic]
            self.h = ctypes.c_void_p(self.lib.plb_b200_create_enveloped(
                q, 8 if prec == "f64" else 4, nx, ny, nz, DYN[dyn], p1, p2, px, py, pz, rank, n_ranks, slabs_per_rank))
        else:
            self.h = ctypes.c_void_p(self.lib.plb_b200_create_distributed(
                q, 8 if prec == "f64" else 4, nx, ny, nz, DYN[dyn], p1, p2, rank, n_ranks, slabs_per_rank))
        if not self.h:
            raise PlbError(self.lib.plb_b200_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.plb_b200_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, ok):
        if not ok:
            raise PlbError(self.lib.plb_b200_last_error().decode())

    @property
    def box(self):
        return (0, self.nx - 1, 0, self.ny - 1, 0, self.nz - 1)

    def comm_init(self, uid: bytes):
        self._ck(self.lib.plb_b200_comm_init(self.h, ctypes.c_char_p(uid)))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.plb_b200_set_stream(self.h, ctypes.c_void_p(int(cuda_stream))))

    def set_periodicity(self, px, py, pz):
        self._ck(self.lib.plb_b200_set_periodicity(self.h, int(px), int(py), int(pz)))

    def toggle_internal_statistics(self, on):
        self._ck(self.lib.plb_b200_toggle_internal_statistics(self.h, int(bool(on))))

    def define_dynamics(self, box, kind, p1, p2=0.0):
        self._ck(self.lib.plb_b200_define_dynamics_box(self.h, _box(box), DYN[kind], p1, p2))

    def define_dynamics_mask(self, mask, flag, kind, p1, p2=0.0):
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        assert mask.shape == (self.nx, self.ny, self.nz)
        self._ck(self.lib.plb_b200_define_dynamics_mask(self.h, mask.ctypes.data_as(ctypes.c_void_p), int(flag), DYN[kind], p1, p2))

    def set_velocity_condition_on_block_boundaries(self, model, block=None, app=None):
        block = block or self.box
        self._ck(self.lib.plb_b200_set_velocity_condition_on_block_boundaries(self.h, BC[model], _box(block), _box(app or block)))

    def add_velocity_boundary(self, model, d, o, box):
        self._ck(self.lib.plb_b200_add_velocity_boundary(self.h, BC[model], d, o, _box(box)))

    def add_pressure_boundary(self, model, d, o, box):
        self._ck(self.lib.plb_b200_add_pressure_boundary(self.h, BC[model], d, o, _box(box)))

    def add_external_velocity_edge(self, model, plane, n1, n2, box):
        self._ck(self.lib.plb_b200_add_external_velocity_edge(self.h, BC[model], plane, n1, n2, _box(box)))

    def add_external_velocity_corner(self, model, xn, yn, zn, x, y, z):
        self._ck(self.lib.plb_b200_add_external_velocity_corner(self.h, BC[model], xn, yn, zn, x, y, z))

    def add_internal_velocity_edge(self, model, plane, n1, n2, box):
        self._ck(self.lib.plb_b200_add_internal_velocity_edge(self.h, BC[model], plane, n1, n2, _box(box)))

    def add_internal_velocity_corner(self, model, xn, yn, zn, x, y, z):
        self._ck(self.lib.plb_b200_add_internal_velocity_corner(self.h, BC[model], xn, yn, zn, x, y, z))

    def set_boundary_velocity(self, box, u):
        self._ck(self.lib.plb_b200_set_boundary_velocity(self.h, _box(box), _vec(u)))

    def set_boundary_density(self, box, rho):
        self._ck(self.lib.plb_b200_set_boundary_density(self.h, _box(box), rho))

    def initialize_at_equilibrium(self, box, rho, u):
        self._ck(self.lib.plb_b200_initialize_at_equilibrium(self.h, _box(box), rho, _vec(u)))

    def initialize_at_equilibrium_field(self, box, rho_u):
        rho_u = np.ascontiguousarray(rho_u, dtype=np.float64)
        self._ck(self.lib.plb_b200_initialize_at_equilibrium_field(self.h, _box(box), rho_u.ctypes.data_as(ctypes.c_void_p)))

    def initialize(self):
        self._ck(self.lib.plb_b200_initialize(self.h))

    def local_cells(self, box=None):
        n = self.lib.plb_b200_local_bytes(self.h, _box(box or self.box))
        return n // (self.q * np.dtype(self.dtype).itemsize)

    def upload(self, pops, box=None):
        """pops: the cells of `box` this rank owns, [cells..., q] in x,y,z order (wire format)."""
        pops = np.ascontiguousarray(pops, dtype=self.dtype)
        self._ck(self.lib.plb_b200_upload(self.h, _box(box or self.box), pops.ctypes.data_as(ctypes.c_void_p), pops.nbytes))

    def upload_ptr(self, ptr, nbytes, box=None):
        self._ck(self.lib.plb_b200_upload(self.h, _box(box or self.box), ctypes.c_void_p(ptr), nbytes))

    def download_ptr(self, ptr, nbytes, box=None):
        self._ck(self.lib.plb_b200_download(self.h, _box(box or self.box), ctypes.c_void_p(ptr), nbytes))

    def download(self, box=None):
        box = box or self.box
        if self.n_ranks == 1:
            shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, self.q)
        else:
            shape = (self.local_cells(box), self.q)
        out = np.empty(shape, dtype=self.dtype)
        self._ck(self.lib.plb_b200_download(self.h, _box(box), out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def collide_and_stream(self, n=1):
        self._ck(self.lib.plb_b200_collide_and_stream(self.h, int(n)))

    def enable_external_rhoj(self):
        """the fields and the two functionals are part of plb_b200_create_enveloped; nothing left to do"""

    def execute_internal_processors(self, n=1):
        self._ck(self.lib.plb_b200_execute_internal_processors(self.h, int(n)))

    def apply_rhobar_j(self, box=None):
        self._ck(self.lib.plb_b200_apply_rhobar_j(self.h, _box(box or self.box)))

    def upload_rhobar_j(self, fld, box=None):
        fld = np.ascontiguousarray(fld, dtype=self.dtype)
        self._ck(self.lib.plb_b200_upload_rhobar_j(self.h, _box(box or self.box), fld.ctypes.data_as(ctypes.c_void_p), fld.nbytes))

    def download_rhobar_j(self, box=None):
        box = box or self.box
        if self.n_ranks == 1:
            shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, 4)
        else:
            shape = (self.local_cells(box), 4)
        out = np.empty(shape, dtype=self.dtype)
        self._ck(self.lib.plb_b200_download_rhobar_j(self.h, _box(box), out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def synchronize(self):
        self._ck(self.lib.plb_b200_synchronize(self.h))

    def stored_statistics(self):
        out = (ctypes.c_double * 3)()
        n = ctypes.c_uint64()
        self._ck(self.lib.plb_b200_get_stored_statistics(self.h, out, ctypes.byref(n)))
        return out[0], out[1], out[2], n.value

    def compute_average_energy(self, box=None):
        out = ctypes.c_double()
        self._ck(self.lib.plb_b200_compute_average_energy(self.h, _box(box or self.box), ctypes.byref(out)))
        return out.value

    def compute_average_density(self, box=None):
        out = ctypes.c_double()
        self._ck(self.lib.plb_b200_compute_average_density(self.h, _box(box or self.box), ctypes.byref(out)))
        return out.value

    def compute_density_velocity(self, box=None):
        box = box or self.box
        if self.n_ranks == 1:
            shape = (box[1] - box[0] + 1, box[3] - box[2] + 1, box[5] - box[4] + 1, 4)
        else:
            shape = (self.local_cells(box), 4)
        out = np.empty(shape, dtype=self.dtype)
        self._ck(self.lib.plb_b200_compute_density_velocity(self.h, _box(box), out.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def launch_count(self):
        return int(self.lib.plb_b200_launch_count(self.h))

    def device_bytes(self):
        return int(self.lib.plb_b200_device_bytes(self.h))

    def profile_sweeps(self, on):
        self._ck(self.lib.plb_b200_profile_sweeps(self.h, int(bool(on))))

    def profile_read(self):
        ms, n = ctypes.c_double(), ctypes.c_int()
        self._ck(self.lib.plb_b200_profile_read(self.h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def halo_transport(self):
        return self.lib.plb_b200_halo_transport(self.h).decode()

    def save_binary_block(self, path):
        """saveBinaryBlock(lattice, fName): checkpoint file in the reference's format"""
        self._ck(self.lib.plb_b200_save_binary_block(self.h, str(path).encode()))

    def parallel_save(self, base_path, data_file_entry=None, global_id=0):
        """parallelIO::save(lattice, fName, dynamicContent=False): <base_path>.plb + <base_path>.dat"""
        entry = None if data_file_entry is None else str(data_file_entry).encode()
        self._ck(self.lib.plb_b200_parallel_save(self.h, str(base_path).encode(), entry, int(global_id)))

    def parallel_load(self, base_path):
        """parallelIO::load(fName, lattice, dynamicContent=False)"""
        self._ck(self.lib.plb_b200_parallel_load(self.h, str(base_path).encode()))

    def load_binary_block(self, path):
        self._ck(self.lib.plb_b200_load_binary_block(self.h, str(path).encode()))

    def boundary_stages(self):
        ns, ni = ctypes.c_int(), ctypes.c_int64()
        self._ck(self.lib.plb_b200_boundary_stages(self.h, ctypes.byref(ns), ctypes.byref(ni)))
        return ns.value, ni.value
