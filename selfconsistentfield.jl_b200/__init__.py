"""scf-b200: B200-native Fock build (J+K) behind SelfConsistentField.jl's
TwoElectronBuilder / add_2e_term! boundary.  See DESIGN.md / INTEGRATION.md."""
from . import capi
from . import distributed, scf, synthetic
from .builder import JKBuilderCUDA, TwoElectronBuilder, comm_unique_id
from .capi import ScfbError
from .dump import dump_molsturm_hdf5

from .scf import (ScfProblem, assemble_hf_problem, break_spin_symmetry, compute_guess,
                  load_integral_hdf5, load_integral_npz, run_scf)

__all__ = ["capi", "scf", "ScfProblem", "assemble_hf_problem", "break_spin_symmetry", "compute_guess",
           "load_integral_hdf5", "load_integral_npz", "run_scf", "dump_molsturm_hdf5", "JKBuilderCUDA", "TwoElectronBuilder", "ScfbError", "comm_unique_id"]
