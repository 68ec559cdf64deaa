#!/usr/bin/env python
"""Regenerate tests/golden/*.npz + golden_energies.json from the reference's fixtures.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

Sources:
  * integral dumps  /root/reference/test/data/integrals_{be,c}_3-21g.hdf5
    (read with the repo's struct+zlib HDF5 parser — no h5py/libhdf5 here);
  * golden energies /root/reference/test/functionality_hartee_fock.jl:18,24,30,36
    (asserted there with rtol=1e-9 together with `converged`).

Arrays are stored exactly as they sit in the file (C order, file dims).  HDF5.jl
hands Julia the same bytes with reversed dims (load_integral_hdf5.jl:11), i.e.
Julia's eri[i,j,k,l] == file[l,k,j,i]; loaders apply `.T` for the Julia view.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "selfconsistentfield.jl_b200"))
from minih5 import MiniH5  # noqa: E402

REF = "/root/reference/test/data"

GOLDEN = {
    # case: (file, restricted, E_total, cite)
    "rhf_be_321g": ("integrals_be_3-21g", True, -14.4868202421763,
                    "test/functionality_hartee_fock.jl:17-19"),
    "uhf_be_321g": ("integrals_be_3-21g", False, -14.4868202421763,
                    "test/functionality_hartee_fock.jl:23-25"),
    "rohf_c_321g": ("integrals_c_3-21g", True, -37.4803888099046,
                    "test/functionality_hartee_fock.jl:29-31"),
    "uhf_c_321g": ("integrals_c_3-21g", False, -37.4810698325847,
                   "test/functionality_hartee_fock.jl:35-37"),
}


def main():
    for stem in ("integrals_be_3-21g", "integrals_c_3-21g"):
        h = MiniH5(os.path.join(REF, stem + ".hdf5"))
        np.savez_compressed(
            os.path.join(HERE, stem + ".npz"),
            electron_repulsion_bbbb=h.read("electron_repulsion_bbbb"),
            kinetic_bb=h.read("kinetic_bb"),
            nuclear_attraction_bb=h.read("nuclear_attraction_bb"),
            overlap_bb=h.read("overlap_bb"),
            nelec=h.read("system/nelec"),
            atom_numbers=h.read("system/atom_numbers"),
            coords=h.read("system/coords"),
            basis_type=np.array(h.read("discretisation/basis_type")),
            basis_set_name=np.array(h.read("discretisation/basis_set_name")),
        )
        print("wrote", stem + ".npz")
    with open(os.path.join(HERE, "golden_energies.json"), "w") as fh:
        json.dump({k: {"file": v[0], "restricted": v[1], "e_total": v[2], "rtol": 1e-9,
                       "cite": v[3]} for k, v in GOLDEN.items()}, fh, indent=1)
    print("wrote golden_energies.json")


if __name__ == "__main__":
    main()
