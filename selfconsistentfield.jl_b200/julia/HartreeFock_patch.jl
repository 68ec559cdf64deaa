# Patch for /root/reference/examples/HartreeFock.jl:67-98 so that the example runs unchanged with a
# JKBuilderCUDA among `problem.terms_2e`.  The reference's helper asserts
# `isa(JKBuilder, JKBuilderFromTensor)` (:84) and contracts `JKBuilder.eri` six more times on the
# host (:88-95); a builder whose tensor lives in HBM has no `.eri`.  The same two scalars are
#     E_J =  1/2 sum_{s,t} tr(D_s J[D_t])      E_K = -1/2 sum_s tr(D_s K[D_s])
# and come from ONE more J/K pass of the CUDA builder (`scfb_jk_split`).
#
# Use: `include("JKBuilderCUDA.jl"); include("HartreeFock_patch.jl")` before `include`-ing the
# example, then replace the body of `compute_termwise_hf_energies!` from the line
# `@assert length(problem.terms_2e) == 1` on by a call of `jk_energy_split!(energies, problem, Dα, Dβ)`:
#
#     -    @assert length(problem.terms_2e) == 1
#     -    JKBuilder = problem.terms_2e["coulomb+exchange"]
#     -    @assert isa(JKBuilder, JKBuilderFromTensor)
#     -    eri = JKBuilder.eri
#     -    @tensor begin ... end
#     -    energies["coulomb"] = energy_J
#     -    energies["exchange"] = energy_K
#     +    jk_energy_split!(energies, problem, Dα, Dβ)
#
# (Python twin, which is what is tested: selfconsistentfield.jl_b200/scf.py
# `compute_termwise_hf_energies`, examples/HartreeFock.py.)
using LinearAlgebra: tr

"Coulomb / exchange energies of the converged densities for any TwoElectronBuilder."
function jk_energy_split!(energies, problem, Dα::AbstractMatrix, Dβ::AbstractMatrix)
    @assert length(problem.terms_2e) == 1
    builder = problem.terms_2e["coulomb+exchange"]
    J, K = jk_matrices(builder, Dα, Dβ)
    energies["coulomb"] = 1/2 * (tr(Dα * J) + tr(Dβ * J))                       # J = J[Dα + Dβ]
    energies["exchange"] = -1/2 * (tr(Dα * K[:, :, 1]) + tr(Dβ * K[:, :, 2]))
    energies
end

# one J/K pass on the GPU
function jk_matrices(builder::JKBuilderCUDA, Dα, Dβ)
    density = cat(Array(Dα), Array(Dβ); dims=3)
    jk_split(builder, density, Array(Dα) + Array(Dβ))
end

# the reference's own builder: same scalars through the plugin boundary (two add_2e_term! calls),
# so the patched example also keeps working with JKBuilderFromTensor
function jk_matrices(builder::TwoElectronBuilder, Dα, Dβ)
    n = size(Dα, 1)
    density = cat(Array(Dα), Array(Dβ); dims=3)
    total = Array(Dα) + Array(Dβ)
    JmK = zeros(n, n, 2)
    add_2e_term!(JmK, builder, density, total)            # J - K_s
    Jonly = zeros(n, n, 2)
    add_2e_term!(Jonly, builder, zeros(n, n, 2), total)   # zero spin densities: K = 0, leaves J
    Jonly[:, :, 1], Jonly .- JmK
end
