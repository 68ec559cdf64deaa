#!/usr/bin/env python
"""Host-side twin of the reference's example driver (BASELINE.json config 1):
    /root/reference/examples/HartreeFock.jl  +  examples/print_results.jl
with the Fock build running on the B200 through libscf_b200.so.

    python examples/HartreeFock.py [--unrestricted|--restricted] <integral_file> [<outfile>]

<integral_file> is an integral dump in the reference's HDF5 schema
(scripts/dump_integrals_pyscf.py:21-51, e.g. test/data/integrals_be_3-21g.hdf5) or one of the
converted fixtures under tests/golden/*.npz.  Same flow as `hartree_fock` (HartreeFock.jl:100-134):
load -> assemble_hf_problem -> hcore guess (-> break_spin_symmetry) -> run_scf with
max_error_norm=5e-7, max_energy_total_change=1.25e-7 -> termwise energies -> print.
<outfile> is written in the molsturm HDF5 layout like the reference does
(src/util/dump_molsturm_hdf5.jl; selfconsistentfield.jl_b200/dump.py), or as a NumPy archive if it ends in .npz.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scf_b200  # noqa: E402
from scf_b200 import scf as S  # noqa: E402


def parse_args(argv):
    """HartreeFock.jl:7-48."""
    restricted, ofile, nxt = None, None, 0
    while nxt < len(argv):
        a = argv[nxt]
        if a == "--unrestricted":
            restricted = False
        elif a == "--restricted":
            restricted = True
        elif a in ("-h", "--help"):
            print(f"{sys.argv[0]} [--unrestricted|--restricted] <integral_file> [<outfile>]")
        else:
            break
        nxt += 1
    if nxt >= len(argv):
        raise SystemExit("No input hdf5 integral file given")
    if not os.path.isfile(argv[nxt]):
        raise SystemExit(f"Not a valid hdf5 integral file: {argv[nxt]}")
    intfile = argv[nxt]
    nxt += 1
    if nxt < len(argv):
        ofile = argv[nxt]
        nxt += 1
    if nxt < len(argv):
        raise SystemExit(f"Excess commandline arguments: {argv[nxt:]}")
    return restricted, intfile, ofile


def print_info(problem):
    """HartreeFock.jl:50-65."""
    scftype = "ROHF"
    if problem.restricted:
        if S.is_closed_shell(problem):
            scftype = "RHF"
    else:
        scftype = "UHF"
    print("Problem information:")
    print("   № α electrons:     ", problem.n_elec[0])
    print("   № β electrons:     ", problem.n_elec[1])
    print("   № basis functions: ", problem.h_core.shape[0])
    print(f"   SCF type:          {scftype}")


def print_results(problem, res):
    """print_results.jl:4-60."""
    print("SCF converged." if res["converged"] else "SCF failed to converge")
    print()
    e = res["energies"]
    print("Final energies")
    for key in ("coulomb", "exchange", "kinetic", "nuclear_attraction", "nuclear_repulsion"):
        print("%20s = %15.10g" % (key, e[key]))
    print()
    print("%20s = %15.10g" % ("E_1e", e["1e"]))
    print("%20s = %15.10g" % ("E_2e", e["2e"]))
    print("%20s = %15.10g" % ("E electronic", e["1e"] + e["2e"]))
    print()
    epot = e["nuclear_attraction"] + e["0e"] + e["2e"]
    print("%20s = %15.10g" % ("E_pot", epot))
    print("%20s = %15.10g" % ("E_kin", e["kinetic"]))
    print("%20s = %15.10g" % ("virial ratio", -epot / e["kinetic"]))
    print()
    print("%20s = %20.15g" % ("E_total", e["total"]))
    print()
    orben = res["orben"]
    n_orb, n_spin = orben.shape
    print("Orbital occupation")
    print("a                             b")
    for i in range(n_orb):
        aocc = "*" if i < problem.n_elec[0] else " "
        bocc = "*" if i < problem.n_elec[1] else " "
        if n_spin == 1:
            print(f"{aocc}       {orben[i, 0]}       {bocc}")
        else:
            print(f"{aocc}       {orben[i, 0]}    |    {orben[i, 1]}       {bocc}")


def hartree_fock(intfile, restricted=None, ofile=None):
    """HartreeFock.jl:100-134."""
    print(f"Reading file {intfile}")
    load = scf_b200.load_integral_npz if intfile.endswith(".npz") else scf_b200.load_integral_hdf5
    system, integrals = load(intfile)
    print()
    problem = scf_b200.assemble_hf_problem(system, integrals, restricted=restricted)
    print_info(problem)
    print()
    guess_density = scf_b200.compute_guess(problem, system["coords"], system["atom_numbers"],
                                           method="hcore")
    if (not problem.restricted) and S.is_closed_shell(problem):
        scf_b200.break_spin_symmetry(guess_density)
    res = scf_b200.run_scf(problem, guess_density, max_error_norm=5e-7,
                           max_energy_total_change=1.25e-7)
    if not res["converged"]:
        raise SystemExit("SCF failed to converge")
    S.compute_termwise_hf_energies(res, problem)
    print_results(problem, res)
    if ofile is not None:
        if ofile.endswith(".npz"):
            np.savez(ofile, orben=res["orben"], orbcoeff=res["orbcoeff"], density=res["density"],
                     fock=res["fock"], **{f"energy_{k}": v for k, v in res["energies"].items()})
        else:                                          # HartreeFock.jl:131-133
            scf_b200.dump_molsturm_hdf5(integrals, res, ofile)
    return res


def main():
    restricted, intfile, ofile = parse_args(sys.argv[1:])
    hartree_fock(intfile, restricted=restricted, ofile=ofile)


if __name__ == "__main__":
    main()
