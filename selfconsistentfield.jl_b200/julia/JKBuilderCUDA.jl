# JKBuilderCUDA.jl — Julia glue for libscf_b200.so (B200-native Fock build).
#
# NOTE: Julia is not installed in the environment this repository is built and tested in, so
# this file has never been executed.  It is deliberately thin: every call below is a 1:1
# `ccall` of a symbol declared in include/scf_b200.h, and the Python ctypes harness
# (selfconsistentfield.jl_b200/{capi,builder,scf}.py), which IS tested on B200s, drives exactly
# the same symbols with the same column-major memory layout.
#
# Drop-in boundary (reference files, relative to SelfConsistentField.jl):
#   src/types/ScfProblem.jl:2-9                  abstract type TwoElectronBuilder, add_2e_term!
#   src/algorithms/JKBuilderFromTensor.jl:7-51   the builder this one replaces
#   src/util/load_integral_hdf5.jl:15,51-53      where the default builder is constructed
#
# Usage:
#   include("JKBuilderCUDA.jl"); using .SCFB200
#   system, integrals = load_integral_hdf5("integrals.hdf5")
#   integrals = with_cuda_builder(integrals)                  # ERI uploaded once, stays in HBM
#   problem = assemble_hf_problem(system, integrals)
#   res = run_scf(problem, compute_guess(problem, system.coords, system.atom_numbers))
module SCFB200

using SelfConsistentField
import SelfConsistentField: TwoElectronBuilder, add_2e_term!

export JKBuilderCUDA, with_cuda_builder, upload_slabs!, mark_ready!, jk_split, run_scf_device, scf_device_setup!,
       scf_device_fock!, scf_device_roothaan!, scf_device_diis_push!, scf_device_diis_extrapolate!

const libscf = get(ENV, "SCFB200_LIB", joinpath(@__DIR__, "..", "libscf_b200.so"))

const SCFB_LAYOUT_FULL = Cint(0)
const SCFB_LAYOUT_PACKED8 = Cint(1)
const SCFB_ERI_HASH = Cint(0)
const SCFB_ERI_CHAIN = Cint(1)

last_error(h::Ptr{Cvoid}) = unsafe_string(ccall((:scfb_last_error, libscf), Cstring, (Ptr{Cvoid},), h))

"Turn a non-zero status into a Julia error (the library never throws across the ABI)."
function check(rc::Cint, h::Ptr{Cvoid}=C_NULL)
    rc == 0 || error("libscf_b200 ($rc): " * last_error(h))
    nothing
end

"""
J+K builder whose ERI tensor lives in B200 HBM (full fp64 layout or 8-fold packed).
`devices=0:7` shards ONE builder over several GPUs of this process (`scfb_create_multi`): the
unmodified, single-threaded `run_scf` then uses all of them, e.g. for an n_bas = 512 tensor.
Otherwise one builder owns the shard of one GPU and `(rank, n_ranks)` select the nu-slab block
(one Julia process per GPU, see `comm_init!`).
"""
mutable struct JKBuilderCUDA <: TwoElectronBuilder
    handle::Ptr{Cvoid}
    n_bas::Int
    layout::Symbol

    function JKBuilderCUDA(n_bas::Integer; layout::Symbol=:full, device::Integer=0,
                           rank::Integer=0, n_ranks::Integer=1, devices=nothing)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        lay = layout == :full ? SCFB_LAYOUT_FULL : SCFB_LAYOUT_PACKED8
        if devices === nothing
            check(ccall((:scfb_create, libscf), Cint,
                        (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint),
                        h, n_bas, lay, device, rank, n_ranks))
        else
            @assert rank == 0 && n_ranks == 1
            devs = Cint[d for d in devices]
            check(ccall((:scfb_create_multi, libscf), Cint,
                        (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Ptr{Cint}),
                        h, n_bas, lay, length(devs), devs))
        end
        obj = new(h[], n_bas, layout)
        finalizer(obj) do b
            b.handle == C_NULL || ccall((:scfb_destroy, libscf), Cint, (Ptr{Cvoid},), b.handle)
            b.handle = C_NULL
        end
        obj
    end
end

"""
Chunked upload for tensors too large for host memory (load_integral_hdf5.jl:9-10 TODO):
`slabs[:,:,:,1:count]` are the nu-slabs `nu_first .. nu_first+count-1` (1-based like every Julia
index; e.g. `h5f["electron_repulsion_bbbb"][:,:,:,r]` for a range `r`).  Slabs outside this
builder's shard are skipped; call `mark_ready!` once the shard is complete.
"""
function upload_slabs!(b::JKBuilderCUDA, slabs::AbstractArray{Float64,4}, nu_first::Integer)
    @assert size(slabs)[1:3] == (b.n_bas, b.n_bas, b.n_bas)
    dense = Array{Float64,4}(slabs)
    check(ccall((:scfb_eri_upload_slabs, libscf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint),
                b.handle, dense, nu_first - 1, size(dense, 4)), b.handle)
    b
end

mark_ready!(b::JKBuilderCUDA) =
    (check(ccall((:scfb_eri_mark_ready, libscf), Cint, (Ptr{Cvoid},), b.handle), b.handle); b)

"Equivalent of `JKBuilderFromTensor(eri)`: upload the dense (n,n,n,n) tensor once."
function JKBuilderCUDA(eri::AbstractArray{Float64,4}; kwargs...)
    n = size(eri, 1)
    @assert size(eri) == (n, n, n, n)
    b = JKBuilderCUDA(n; kwargs...)
    dense = Array{Float64,4}(eri)       # column-major, exactly the layout the library expects
    check(ccall((:scfb_eri_upload, libscf), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.handle, dense),
          b.handle)
    b
end

"ERI generated in HBM from a closed-form family (`:hash` or `:chain`), see DESIGN.md."
function JKBuilderCUDA(n_bas::Integer, family::Symbol, seed::Integer;
                       aux::Union{Nothing,Vector{Float64}}=nothing, kwargs...)
    b = JKBuilderCUDA(n_bas; kwargs...)
    fam = family == :hash ? SCFB_ERI_HASH : SCFB_ERI_CHAIN
    auxp = aux === nothing ? Ptr{Float64}(C_NULL) : pointer(aux)
    GC.@preserve aux check(ccall((:scfb_eri_generate, libscf), Cint,
                                 (Ptr{Cvoid}, Cint, UInt64, Ptr{Float64}),
                                 b.handle, fam, UInt64(seed), auxp), b.handle)
    b
end

"""
`out[:,:,s] += J[total_density] - K[density[:,:,s]]` — the method the SCF loop reaches through
`compute_fock_matrix` (src/parts/compute_fock_matrix.jl:31-34).  Same assertions as the
reference method (src/algorithms/JKBuilderFromTensor.jl:30-32).
"""
function add_2e_term!(out::AbstractArray, builder::JKBuilderCUDA,
                      density::AbstractArray, total_density::AbstractArray)
    n_bas, _, n_spin = size(density)
    @assert n_spin == 1 || n_spin == 2
    @assert size(total_density) == (n_bas, n_bas)
    @assert size(density) == (n_bas, n_bas, n_spin)
    @assert n_bas == builder.n_bas
    # `convert` returns a dense Array{Float64} argument itself and only copies views / other eltypes
    # (the `Array{Float64,3}(x)` constructor would copy on every Fock build)
    d = convert(Array{Float64,3}, density)
    t = convert(Array{Float64,2}, total_density)
    o = out isa Array{Float64,3} ? out : Array{Float64,3}(out)
    check(ccall((:scfb_fock_jk, libscf), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                builder.handle, n_spin, d, t, o), builder.handle)
    o === out || (out .= o)
    out
end

"J (n,n) and K (n,n,n_spin) separately — for the post-SCF energy split of examples/HartreeFock.jl:88-95."
function jk_split(builder::JKBuilderCUDA, density::AbstractArray, total_density::AbstractArray)
    n, _, n_spin = size(density)
    J = zeros(n, n); K = zeros(n, n, n_spin)
    check(ccall((:scfb_jk_split, libscf), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                builder.handle, n_spin, Array{Float64,3}(density), Array{Float64,2}(total_density),
                J, K), builder.handle)
    J, K
end

"Replace the default builder of `load_integral_hdf5`'s result by a CUDA one."
function with_cuda_builder(integrals; kwargs...)
    merge(integrals, (jk_builder_bb=JKBuilderCUDA(integrals.electron_repulsion_bbbb; kwargs...),))
end

# ---------------------------------------------------------------------------------------
# Multi-GPU (one Julia process per GPU, e.g. under MPI.jl): rank 0 creates the NCCL id, every
# rank joins; afterwards add_2e_term! returns the allreduced result on every rank.
# ---------------------------------------------------------------------------------------
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:scfb_comm_unique_id, libscf), Cint, (Ptr{UInt8},), id))
    id                                        # ship to the other ranks, e.g. MPI.Bcast!(id, 0, comm)
end

comm_init!(b::JKBuilderCUDA, id::Vector{UInt8}) =
    check(ccall((:scfb_comm_init, libscf), Cint, (Ptr{Cvoid}, Ptr{UInt8}), b.handle, id), b.handle)

# ---------------------------------------------------------------------------------------
# Device-resident SCF step (scfb_scf_*): S, h_core, D, F, error, orbitals and the DIIS history
# stay in HBM; only scalars, DIIS rows and coefficients cross PCIe.  The caller keeps the
# control flow of src/run_scf.jl:21-49 and the small solves of cDIIS.jl:33-63 / EDIIS.jl:20-44.
# ---------------------------------------------------------------------------------------
function scf_device_setup!(b::JKBuilderCUDA, problem::ScfProblem; n_diis_size::Integer=5)
    e0 = Float64(sum(v for (k, v) in problem.terms_0e))
    check(ccall((:scfb_scf_setup, libscf), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cint, Cint, Cint, Cint, Cint),
                b.handle, Array{Float64,2}(problem.overlap), Array{Float64,2}(problem.h_core), e0,
                problem.n_elec[1], problem.n_elec[2], problem.n_orb, problem.restricted ? 1 : 0,
                n_diis_size), b.handle)
end

function scf_device_set_density!(b::JKBuilderCUDA, density::AbstractArray)
    check(ccall((:scfb_scf_set_density, libscf), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}),
                b.handle, size(density, 3), Array{Float64,3}(density)), b.handle)
end

"compute_fock_matrix on the resident density -> (energies::Dict, error_norm)."
function scf_device_fock!(b::JKBuilderCUDA)
    e = zeros(4); nrm = Ref{Cdouble}(0)
    check(ccall((:scfb_scf_fock, libscf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Cdouble}),
                b.handle, e, nrm), b.handle)
    Dict("0e" => e[1], "1e" => e[2], "2e" => e[3], "total" => e[4]), nrm[]
end

"compute_orbitals + compute_density on the resident Fock iterate (Roothaan.jl:5-6)."
scf_device_roothaan!(b::JKBuilderCUDA) =
    check(ccall((:scfb_scf_roothaan, libscf), Cint, (Ptr{Cvoid},), b.handle), b.handle)

"Push the newest (F, e, D) of `spin` (1-based) and return the new cDIIS / EDIIS matrix rows."
function scf_device_diis_push!(b::JKBuilderCUDA, spin::Integer)
    m = Ref{Cint}(0); rc = zeros(16); re = zeros(16)
    check(ccall((:scfb_scf_diis_push, libscf), Cint,
                (Ptr{Cvoid}, Cint, Ref{Cint}, Ptr{Float64}, Ptr{Float64}),
                b.handle, spin - 1, m, rc, re), b.handle)
    rc[1:m[]], re[1:m[]]
end

"F[:,:,spin] <- sum_i c[i] F_i over the device history, newest first (parts/diis.jl:52-54)."
scf_device_diis_extrapolate!(b::JKBuilderCUDA, spin::Integer, c::Vector{Float64}) =
    check(ccall((:scfb_scf_diis_extrapolate, libscf), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}),
                b.handle, spin - 1, length(c), c), b.handle)

scf_device_diis_purge!(b::JKBuilderCUDA, spin::Integer, count::Integer) =
    check(ccall((:scfb_scf_diis_purge, libscf), Cint, (Ptr{Cvoid}, Cint, Cint),
                b.handle, spin - 1, count), b.handle)

scf_device_diis_reset!(b::JKBuilderCUDA) =
    check(ccall((:scfb_scf_diis_reset, libscf), Cint, (Ptr{Cvoid},), b.handle), b.handle)

"Fetch a resident quantity: 0 fock, 1 fock_new, 2 density, 3 error, 4 orbcoeff, 5 orben (+_PREV 6,7,8)."
function scf_device_get(b::JKBuilderCUDA, what::Integer, n_spin::Integer)
    n = b.n_bas
    out = what in (5, 7) ? zeros(n, n_spin) : zeros(n, n, n_spin)
    check(ccall((:scfb_scf_get, libscf), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}),
                b.handle, what, out), b.handle)
    out
end

# ---------------------------------------------------------------------------------------
# run_scf with the whole iteration on the device.  Control flow, defaults and the result Dict
# follow src/run_scf.jl:4-65 line by line; the small coefficient solves are the reference's
# own `diis_solve_coefficients` methods (cDIIS.jl:33-63, EDIIS.jl:20-44), fed with matrices
# assembled from the rows the device returns.  (Python twin, which is what is tested:
# selfconsistentfield.jl_b200/scf.py `_run_scf_device`.)
# ---------------------------------------------------------------------------------------
using DataStructures: CircularBuffer
import SelfConsistentField: cDIIS, EDIIS, diis_solve_coefficients, check_convergence,
                            ScfConvergence

"Insert the new first row/column of a DIIS matrix (newest first) and drop the oldest beyond `cap`."
function push_row(B::Matrix{Float64}, row::Vector{Float64}, cap::Int)
    m = length(row)
    Bn = zeros(m, m)
    Bn[1, :] = row; Bn[:, 1] = row
    k = m - 1
    k > 0 && (Bn[2:end, 2:end] = B[1:k, 1:k])
    Bn
end

function run_scf_device(problem::ScfProblem, guess_density::AbstractArray;
                        max_iter=100, ediis_max_error_norm=0.1, n_diis_size=5,
                        conditioning_threshold=1e-14, coefficient_threshold=1e-6, kwargs...)
    @assert length(problem.terms_2e) == 1
    b = first(values(problem.terms_2e))::JKBuilderCUDA
    scf_device_setup!(b, problem; n_diis_size=n_diis_size)
    # spin blocks of the Fock iterate: 1 for RHF and ROHF (two densities, one effective Fock block,
    # compute_fock_matrix.jl:61-65), 2 for UHF — what spinloop(iterstate.fock) runs over (diis.jl:8)
    ns = Ref{Cint}(0)
    check(ccall((:scfb_scf_n_spin, libscf), Cint, (Ptr{Cvoid}, Ref{Cint}), b.handle, ns), b.handle)
    n_spin = Int(ns[])
    scf_device_set_density!(b, guess_density)
    energies, _ = scf_device_fock!(b)                       # run_scf.jl:16
    use_ediis = true                                        # run_scf.jl:8
    B = [zeros(0, 0) for _ in 1:n_spin]
    energybuffer = CircularBuffer{Dict}(n_diis_size)
    scfconv = nothing; n_iter = 0; newenergies = energies
    for i in 1:max_iter
        n_iter = i
        # --- compute_next_iterate (parts/diis.jl:4-36) ---
        pushfirst!(energybuffer, energies)
        for σ in 1:n_spin
            rc, re = scf_device_diis_push!(b, σ)
            B[σ] = push_row(B[σ], use_ediis ? re : rc, n_diis_size)
        end
        acc = use_ediis ? EDIIS(problem; n_diis_size=n_diis_size) : cDIIS(problem; n_diis_size=n_diis_size)
        system(σ) = use_ediis ? B[σ] : [B[σ] ones(size(B[σ], 1)); ones(size(B[σ], 1))' 0.0]
        # sync_spins = true: EDIIS adds the spin blocks (EDIIS.jl:126-128), cDIIS keeps the
        # alpha block (cDIIS.jl:69-72)
        A = (n_spin == 2 && use_ediis) ? system(1) + system(2) : system(1)
        c, n_purge = diis_solve_coefficients(acc, A; conditioning_threshold=conditioning_threshold,
                                             energybuffer=energybuffer)
        for σ in 1:n_spin
            mask = map(x -> abs(x) >= coefficient_threshold, c); mask[1] = true   # diis.jl:45-47
            c[argmax(c)] += sum(c[.!mask])
            scf_device_diis_extrapolate!(b, σ, Vector{Float64}(c .* mask))
        end
        for σ in 1:n_spin
            scf_device_diis_purge!(b, σ, n_purge)
            n_purge > 0 && (m = max(0, size(B[σ], 1) - (n_purge + 1)); B[σ] = B[σ][1:m, 1:m])
        end
        # --- roothan_step (Roothaan.jl:5-7) ---
        scf_device_roothaan!(b)
        newenergies, error_norm = scf_device_fock!(b)
        # --- check_convergence (check_convergence.jl:13-37), signed changes ---
        change = Dict(k => newenergies[k] - v for (k, v) in energies)
        converged = !(error_norm >= get(kwargs, :max_error_norm, 5e-7) ||
                      change["total"] >= get(kwargs, :max_energy_total_change, 1.25e-7) ||
                      change["1e"] >= get(kwargs, :max_energy_1e_change, 5e-5))
        scfconv = ScfConvergence(error_norm, change, converged)
        converged && break
        if error_norm < ediis_max_error_norm && use_ediis      # run_scf.jl:42-46
            use_ediis = false
            scf_device_diis_reset!(b)
            B = [zeros(0, 0) for _ in 1:n_spin]
        end
        energies = newenergies
    end
    stale = scfconv.is_converged          # run_scf.jl:40 breaks before :48
    Dict("n_iter" => n_iter, "n_applies" => NaN, "problem" => problem,
         "orben" => scf_device_get(b, stale ? 7 : 5, n_spin)[1:problem.n_orb, :],
         "orbcoeff" => scf_device_get(b, stale ? 6 : 4, n_spin)[:, 1:problem.n_orb, :],
         "density" => scf_device_get(b, stale ? 8 : 2, n_spin),
         "converged" => scfconv.is_converged,
         "fock" => scf_device_get(b, stale ? 0 : 1, n_spin),
         "energies" => stale ? energies : newenergies,
         "error_norm" => scfconv.error_norm, "energy_change" => scfconv.energy_change)
end

end # module
