# B200Backend.jl -- Julia glue a maintainer adds to GaussianMarkovRandomFields.jl (e.g. as a package extension
# `ext/GaussianMarkovRandomFieldsB200.jl`) to put libgmrf_b200.so behind the existing `WorkspaceBackend` protocol
# (src/workspace/backend.jl:8-30). It mirrors `CliqueTreesBackend` (src/workspace/cliquetrees_backend.jl:21-150)
# method by method; everything above it (GMRFWorkspace, WorkspaceGMRF, gaussian_approximation, autodiff rules) runs
# unchanged.  NOT executed in this repository's CI: the image has no Julia. The Python class
# gmrf_b200/backend.py binds exactly the same entry points and is what the parity tests drive.

module GaussianMarkovRandomFieldsB200

using GaussianMarkovRandomFields
using LinearAlgebra, SparseArrays
import GaussianMarkovRandomFields: WorkspaceBackend, GMRFWorkspace, refactorize!, backend_solve, compute_logdet,
    compute_selinv!, get_selinv, get_selinv_diag, backend_backward_solve, selinv_dot, selinv_extract_at,
    ordering_permutation, AbstractLatentWorkspacePool, checkout, checkin, CholeskySqrt
using LinearMaps: LinearMap

const libgmrf = get(ENV, "GMRF_B200_LIB", "libgmrf_b200.so")

export B200Backend, B200WorkspacePool

mutable struct B200Backend <: WorkspaceBackend
    handle::Ptr{Cvoid}
    n::Int
    nnz::Int
    device::Int
    check::Bool
    # field names mirrored from CHOLMODBackend because tests/benchmarks peek at them
    # (test/workspace/test_gmrf_workspace.jl:214,219; benchmarks/benchmarks.jl:185-190)
    selinv_cache::Union{Nothing, SparseMatrixCSC{Float64, Int}}
    selinv_diag_cache::Union{Nothing, Vector{Float64}}
    selinv_pattern::Union{Nothing, Tuple{Vector{Int}, Vector{Int}}}
    factor_pattern::Union{Nothing, Tuple{Vector{Int}, Vector{Int}}}
end

_errmsg(h) = unsafe_string(ccall((:gmrf_b200_last_error, libgmrf), Cstring, (Ptr{Cvoid},), h))

function _check(b::B200Backend, rc::Cint)
    rc == 0 && return nothing
    rc == -1 && throw(ArgumentError(_errmsg(b.handle)))
    rc > 0 && (b.check ? throw(PosDefException(Int(rc))) : return nothing)   # reference uses check=false (backend.jl:184)
    error("libgmrf_b200 error $rc: " * _errmsg(b.handle))
end

"""
    B200Backend(Q::SparseMatrixCSC{Float64,Int}; ordering = nothing, device = 0, check = false)

Symbolic analysis once (host) + first numeric factorization on `device`. `ordering` accepts what
`CHOLMODBackend` accepts (backend.jl:147-153): `nothing` (library default: nested dissection), a permutation
vector, a CliqueTrees elimination algorithm, or `PinDenseColumns(...)` -- resolved on the host by the package's
own `ordering_permutation` and handed over as a 1-based permutation.
"""
function B200Backend(Q::SparseMatrixCSC{Float64, Int}; ordering = nothing, device::Integer = 0, check::Bool = false,
        analysis::Union{Nothing, Vector{UInt8}} = nothing)
    n = size(Q, 1)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    if analysis !== nothing
        # symbolic analysis read from `export_analysis(other_backend)` (same pattern): a later session on the same mesh,
        # or the other GPUs of a pool -- ordering, elimination tree, supernodes and schedules are not recomputed
        rc = GC.@preserve Q analysis ccall((:gmrf_b200_create_from_analysis, libgmrf), Cint,
            (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{UInt8}, Int64, Cint),
            href, n, SparseArrays.getcolptr(Q), rowvals(Q), 1, analysis, length(analysis), device)
    else
    permvec = ordering === nothing ? Int[] : ordering_permutation(Q, ordering)
    rc = GC.@preserve Q permvec ccall((:gmrf_b200_create, libgmrf), Cint,
        (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Int64}, Cint, Cint),
        href, n, SparseArrays.getcolptr(Q), rowvals(Q), 1, isempty(permvec) ? C_NULL : pointer(permvec), 1, device)
    end
    rc == 0 || throw(ArgumentError(unsafe_string(ccall((:gmrf_b200_last_error, libgmrf), Cstring, (Ptr{Cvoid},), C_NULL))))
    b = B200Backend(href[], n, nnz(Q), device, check, nothing, nothing, nothing, nothing)
    finalizer(x -> ccall((:gmrf_b200_destroy, libgmrf), Cvoid, (Ptr{Cvoid},), x.handle), b)
    refactorize!(b, Symmetric(Q))
    return b
end

function refactorize!(b::B200Backend, Q::Symmetric)
    nz = nonzeros(Q.data)                                    # positional, same pattern (gmrf_workspace.jl:131-143)
    _check(b, GC.@preserve nz ccall((:gmrf_b200_refactorize, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64),
        b.handle, nz, length(nz)))
    b.selinv_cache = nothing                                 # backend.jl:185-187
    b.selinv_diag_cache = nothing
    return nothing
end

function backend_solve(b::B200Backend, rhs::AbstractVector)
    B = Vector{Float64}(rhs); X = similar(B)
    _check(b, ccall((:gmrf_b200_solve, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64),
        b.handle, B, X, b.n, 1))
    return X
end

function backend_solve(b::B200Backend, RHS::Matrix{Float64})   # blocked multi-RHS (backend.jl:207-209)
    X = similar(RHS)
    _check(b, ccall((:gmrf_b200_solve, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64),
        b.handle, RHS, X, b.n, size(RHS, 2)))
    return X
end

function backend_backward_solve(b::B200Backend, x::AbstractVector)   # factor.UP \ x (backend.jl:281-284)
    Z = Vector{Float64}(x); X = similar(Z)
    _check(b, ccall((:gmrf_b200_solve_Lt, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64),
        b.handle, Z, X, b.n, 1))
    return X
end

# blocked half solve: column i of Z is the i-th randn! draw, so a matrix `_rand!` keeps the reference's random stream
function backend_backward_solve(b::B200Backend, Z::Matrix{Float64})
    X = similar(Z)
    _check(b, ccall((:gmrf_b200_solve_Lt, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64),
        b.handle, Z, X, b.n, size(Z, 2)))
    return X
end

function compute_logdet(b::B200Backend)
    out = Ref{Float64}(0.0)
    _check(b, ccall((:gmrf_b200_logdet, libgmrf), Cint, (Ptr{Cvoid}, Ref{Float64}), b.handle, out))
    return out[]
end

compute_selinv!(b::B200Backend) = nothing                      # lazy, like backend.jl:215-221

function get_selinv_diag(b::B200Backend)
    if b.selinv_diag_cache === nothing
        if b.selinv_cache !== nothing
            b.selinv_diag_cache = diag(b.selinv_cache)
        else
            d = Vector{Float64}(undef, b.n)
            _check(b, ccall((:gmrf_b200_selinv_diag, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.handle, d))
            b.selinv_diag_cache = d
        end
    end
    return b.selinv_diag_cache
end

function get_selinv(b::B200Backend)
    if b.selinv_cache === nothing
        if b.selinv_pattern === nothing
            nz = Ref{Int64}(0)
            _check(b, ccall((:gmrf_b200_selinv_nnz, libgmrf), Cint, (Ptr{Cvoid}, Ref{Int64}), b.handle, nz))
            cp = Vector{Int}(undef, b.n + 1); rv = Vector{Int}(undef, nz[])
            _check(b, ccall((:gmrf_b200_selinv_pattern, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Cint),
                b.handle, cp, rv, 1))
            b.selinv_pattern = (cp, rv)
        end
        cp, rv = b.selinv_pattern
        vals = Vector{Float64}(undef, length(rv))
        _check(b, ccall((:gmrf_b200_selinv_values, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.handle, vals))
        b.selinv_cache = SparseMatrixCSC(b.n, b.n, copy(cp), copy(rv), vals)
    end
    return b.selinv_cache
end

function selinv_extract_at(b::B200Backend, B::SparseMatrixCSC)
    out = Vector{Float64}(undef, nnz(B))
    cp = Vector{Int64}(SparseArrays.getcolptr(B)); rv = Vector{Int64}(rowvals(B))
    _check(b, ccall((:gmrf_b200_selinv_extract, libgmrf), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Float64}), b.handle, b.n, cp, rv, 1, out))
    return SparseMatrixCSC(size(B)..., copy(SparseArrays.getcolptr(B)), copy(rowvals(B)), out)
end

# tr(Q^-1 B). Float64 B: gathered and contracted on the device, one double comes back. Any other element type
# (ForwardDiff.Dual-valued B, ext/forwarddiff/logdetcov.jl:23): values are read on the device at B's pattern and the
# O(nnz(B)) dot stays generic.
selinv_dot(b::B200Backend, B::SparseMatrixCSC) = dot(nonzeros(selinv_extract_at(b, B)), nonzeros(B))
function selinv_dot(b::B200Backend, B::SparseMatrixCSC{Float64})
    out = Ref{Float64}(0.0)
    cp = Vector{Int64}(SparseArrays.getcolptr(B)); rv = Vector{Int64}(rowvals(B))
    _check(b, ccall((:gmrf_b200_selinv_dot, libgmrf), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Float64}, Ref{Float64}), b.handle, b.n, cp, rv, 1, nonzeros(B), out))
    return out[]
end

# The square root P'L of Q as a sparse matrix (sparse_cho_sqrt, src/linear_maps/cholesky_sqrt.jl:6-21): what
# `CholeskySqrt(cho)` wraps for a CHOLMOD factor. Pattern once per backend, values gathered on the device.
function cholesky_sqrt(b::B200Backend)
    if b.factor_pattern === nothing
        nz = Ref{Int64}(0)
        _check(b, ccall((:gmrf_b200_factor_nnz, libgmrf), Cint, (Ptr{Cvoid}, Ref{Int64}), b.handle, nz))
        cp = Vector{Int64}(undef, b.n + 1); rv = Vector{Int64}(undef, nz[])
        _check(b, ccall((:gmrf_b200_factor_pattern, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Cint), b.handle, cp, rv, 1))
        b.factor_pattern = (cp, rv)
    end
    cp, rv = b.factor_pattern
    vals = Vector{Float64}(undef, length(rv))
    _check(b, ccall((:gmrf_b200_factor_values, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.handle, vals))
    return SparseMatrixCSC(b.n, b.n, copy(cp), copy(rv), vals)
end
CholeskySqrt(b::B200Backend) = LinearMap(cholesky_sqrt(b))

# The symbolic analysis as bytes (write it next to the mesh; `B200Backend(Q; analysis = bytes)` restores it)
function export_analysis(b::B200Backend)
    nb = Ref{Int64}(0)
    _check(b, ccall((:gmrf_b200_analysis_export, libgmrf), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64, Ref{Int64}), b.handle, C_NULL, 0, nb))
    buf = Vector{UInt8}(undef, nb[])
    _check(b, ccall((:gmrf_b200_analysis_export, libgmrf), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64, Ref{Int64}), b.handle, buf, length(buf), nb))
    return buf
end

# GMRFWorkspace(Q, B200Backend; ...) -- copy of cliquetrees_backend.jl:132-150
function GMRFWorkspace(Q::SparseMatrixCSC{T}, ::Type{B200Backend}; kw...) where {T}
    n = size(Q, 1)
    size(Q, 1) == size(Q, 2) || throw(ArgumentError("Q must be square"))
    backend = B200Backend(SparseMatrixCSC{Float64, Int}(Q); kw...)
    return GMRFWorkspace{T, typeof(backend)}(copy(Q), backend, zeros(T, n), zeros(T, n), true, false, false, zero(T), 1, 0)
end

# One workspace per GPU, ordering resolved once (workspace_pool.jl:53-67); protocol of src/workspace/workspace.jl:25-41
struct B200WorkspacePool <: AbstractLatentWorkspacePool
    channel::Channel{Any}
end
function B200WorkspacePool(Q::SparseMatrixCSC; devices = 0:0, ordering = nothing)
    perm = ordering === nothing ? nothing : ordering_permutation(Q, ordering)
    ch = Channel{Any}(length(devices))
    blob = nothing                                   # one analysis for the whole pool (workspace_pool.jl:55-58)
    for d in devices
        ws = blob === nothing ? GMRFWorkspace(Q, B200Backend; ordering = perm, device = d) :
            GMRFWorkspace(Q, B200Backend; analysis = blob, device = d)
        blob === nothing && (blob = export_analysis(ws.backend))
        put!(ch, ws)
    end
    return B200WorkspacePool(ch)
end
checkout(p::B200WorkspacePool) = take!(p.channel)
checkin(p::B200WorkspacePool, ws) = put!(p.channel, ws)

# ---- hyperparameter loops without re-uploading nzval --------------------------------------------------------------
# A model whose precision is a fixed linear combination of value arrays on the workspace pattern (the Matern SPDE:
# K C^-1 K ... expands binomially in kappa^2; fem_utils.jl:313-335 lays every term out on the structural pattern)
# uploads the arrays once and afterwards refactorizes from `length(coeff)` doubles.
function set_value_basis!(b::B200Backend, basis::Matrix{Float64})          # nnz x nbasis, column j = j-th value array
    size(basis, 1) == b.nnz || throw(ArgumentError("basis columns must hold nnz(Q) values"))
    _check(b, ccall((:gmrf_b200_set_value_basis, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), b.handle, basis, size(basis, 2)))
end
function refactorize_combination!(b::B200Backend, coeff::Vector{Float64})
    _check(b, ccall((:gmrf_b200_refactorize_combination, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), b.handle, coeff, length(coeff)))
    b.selinv_cache = nothing; b.selinv_diag_cache = nothing
    return nothing
end

# tr(Q^-1 B_j) for every array of the resident value basis = d logdet Q / d c_j: what the logdetcov / logpdf pullbacks
# (src/workspace/autodiff.jl:8-91) contract Q-bar with when Q(theta) = sum_j c_j(theta) B_j; nbasis doubles come back.
function selinv_dot_basis(b::B200Backend, nbasis::Integer)
    out = Vector{Float64}(undef, nbasis)
    _check(b, ccall((:gmrf_b200_selinv_dot_basis, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), b.handle, out, nbasis))
    return out
end

# Newton loop, diagonal observation Hessian: the iterate Q_prior - H is formed in HBM from the resident prior values.
set_base_values!(b::B200Backend, nzval::Vector{Float64}) =
    _check(b, ccall((:gmrf_b200_set_base_values, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), b.handle, nzval, length(nzval)))
function refactorize_minus_diag!(b::B200Backend, hdiag::Vector{Float64})
    _check(b, ccall((:gmrf_b200_refactorize_base_minus_diag, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), b.handle, hdiag, length(hdiag)))
    b.selinv_cache = nothing; b.selinv_diag_cache = nothing
    return nothing
end

# ... and a SPARSE observation Hessian: `_sparse_hessian_map` (src/workspace/gaussian_approximation.jl:31-61) is uploaded once per
# Newton loop as 1-based nzval positions, every iterate moves nnz(H) values (`_subtract_sparse_hessian!` :74-83 in HBM).
set_hessian_pattern!(b::B200Backend, nzpos::Vector{Int}) =
    _check(b, ccall((:gmrf_b200_set_hessian_pattern, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Cint), b.handle, nzpos, length(nzpos), 1))
function refactorize_minus_sparse!(b::B200Backend, hvals::Vector{Float64})
    _check(b, ccall((:gmrf_b200_refactorize_base_minus_sparse, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), b.handle, hvals, length(hvals)))
    b.selinv_cache = nothing; b.selinv_diag_cache = nothing
    return nothing
end

# diag(A Sigma A') of a sparse design matrix, contracted on the device: the workspace method of `_row_diag_AΣAt`
# (src/linear_predictor_marginals.jl:137-141). A is handed over as CSR = the CSC arrays of its transpose.
function row_diag_AΣAt(b::B200Backend, A::SparseMatrixCSC{Float64,Int})
    At = sparse(transpose(A))
    out = Vector{Float64}(undef, size(A, 1))
    _check(b, GC.@preserve At ccall((:gmrf_b200_selinv_quadform_rows, libgmrf), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Ptr{Float64}),
        b.handle, size(A, 1), SparseArrays.getcolptr(At), rowvals(At), nonzeros(At), 1, out))
    return out
end

# Lanes: B value sets of the same pattern per launch (handle created after `set_option("lanes", B)`); returns the B
# log-determinants and status words. Lane 0 stays the backend's factor.
set_option(key::AbstractString, value::Real) = ccall((:gmrf_b200_set_option, libgmrf), Cint, (Cstring, Cdouble), key, value)
function refactorize_lanes!(b::B200Backend, nzvals::Matrix{Float64})       # nnz x B
    size(nzvals, 1) == b.nnz || throw(ArgumentError("each column must hold nnz(Q) values"))
    B = size(nzvals, 2); ld = Vector{Float64}(undef, B); st = zeros(Cint, B)
    _check(b, ccall((:gmrf_b200_refactorize_lanes, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ptr{Cint}),
                    b.handle, nzvals, b.nnz, B, ld, st))
    b.selinv_cache = nothing; b.selinv_diag_cache = nothing
    return ld, st
end
function refactorize_combination_lanes!(b::B200Backend, coeff::Matrix{Float64})   # nbasis x B
    B = size(coeff, 2); ld = Vector{Float64}(undef, B); st = zeros(Cint, B)
    _check(b, ccall((:gmrf_b200_refactorize_combination_lanes, libgmrf), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Cint}),
                    b.handle, coeff, size(coeff, 1), B, ld, st))
    b.selinv_cache = nothing; b.selinv_diag_cache = nothing
    return ld, st
end

# ---- one factorization, right-hand-side blocks on every GPU (SURVEY.md 8e, config 5) -------------------------------
# `device_array` exposes the numeric factor in HBM (which = 0 panels of L, 1 inverted diagonal blocks, 2 Z panels) so the
# pool can move it between its handles (NCCL.jl `Broadcast!` on `unsafe_wrap(CuArray, ...)`, or a peer copy);
# the receiver then declares it its own with `adopt_factor!` and serves `backend_solve` / `backend_backward_solve`
# for its block of columns without factorizing.
function device_array(b::B200Backend, which::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL); n = Ref{Int64}(0)
    _check(b, ccall((:gmrf_b200_device_array, libgmrf), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}, Ref{Int64}), b.handle, which, p, n))
    return p[], n[]
end
adopt_factor!(b::B200Backend, logdet::Float64; with_selinv::Bool = false) =
    _check(b, ccall((:gmrf_b200_adopt_factor, libgmrf), Cint, (Ptr{Cvoid}, Cdouble, Cint), b.handle, logdet, with_selinv))
# ... guarded: the sender's analysis fingerprint must equal the receiver's, its pivot status travels along
function analysis_fingerprint(b::B200Backend)
    f = Ref{UInt64}(0)
    _check(b, ccall((:gmrf_b200_analysis_fingerprint, libgmrf), Cint, (Ptr{Cvoid}, Ref{UInt64}), b.handle, f))
    return f[]
end
adopt_factor!(b::B200Backend, fingerprint::UInt64, logdet::Float64, status::Integer; with_selinv::Bool = false) =
    _check(b, ccall((:gmrf_b200_adopt_factor_checked, libgmrf), Cint, (Ptr{Cvoid}, UInt64, Cdouble, Cint, Cint),
                    b.handle, fingerprint, logdet, status, with_selinv))

end # module
