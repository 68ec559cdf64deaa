"""scf-b200: B200-native Fock build (J+K) behind SelfConsistentField.jl's
TwoElectronBuilder / add_2e_term! boundary.  See DESIGN.md / INTEGRATION.md."""
from . import capi
from . import distributed, synthetic
from .builder import JKBuilderCUDA, TwoElectronBuilder, comm_unique_id
from .capi import ScfbError

__all__ = ["capi", "JKBuilderCUDA", "TwoElectronBuilder", "ScfbError", "comm_unique_id"]
