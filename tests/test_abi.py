"""The C-ABI library loads without a GPU and exports every symbol include/plb_b200.h declares;
without a device it refuses to create a lattice (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import palabos_b200 as pb
    lib = pb.lib()
    header = open(os.path.join(ROOT, "include", "plb_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(plb_b200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(pb.SYMBOLS) == declared
    assert lib.plb_b200_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import palabos_b200 as pb
    with pytest.raises(pb.PlbError, match="no CUDA device"):
        pb.Lattice(19, "f64", 8, 8, 8, "BGK", 1.0)
    with pytest.raises(pb.PlbError, match="no CUDA device"):
        pb.Lattice(19, "f64", 8, 8, 8, "INCBGK", 1.0, enveloped_periodic=(0, 1, 0))
    h = ctypes.c_int(-1)
    assert pb.lib().plb_b200_coproc_add_domain(8, 4, 4, 4, 1.0, ctypes.byref(h)) == 0


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under palabos_b200/ or include/ may mention it."""
    for base in ("palabos_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath:
                continue
            for fn in files:
                if fn.endswith((".so", ".o", ".pyc")):
                    continue
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "liboracle" not in text and "plbo_" not in text and "oracle/" not in text, os.path.join(dirpath, fn)
