# B200LinearSolve.jl -- boundary B: a LinearSolve.jl algorithm over libgmrf_b200.so so that plain
# `GMRF(mu, Q, B200CholeskyFactorization())`, `linear_condition` (src/arithmetic/condition/linear.jl:46-64) and the
# cache-backed Newton loop (src/arithmetic/condition/gaussian_approximation.jl:428-498) land on the GPU as well.
# It follows ext/GaussianMarkovRandomFieldsPardiso.jl:10-80 method for method: the algorithm object carries the options,
# the LinearSolve cacheval is a `B200Backend`, and the package's dispatch tables get one method each.
# NOT executed in this repository (no Julia in the image); gmrf_b200/gmrf.py is the executable mirror of the same path.

module GaussianMarkovRandomFieldsB200LinearSolve

using GaussianMarkovRandomFields, LinearSolve, LinearAlgebra, SparseArrays
using ..GaussianMarkovRandomFieldsB200: B200Backend
import GaussianMarkovRandomFields: refactorize!, backend_solve, backend_backward_solve, compute_logdet,
    get_selinv, get_selinv_diag

export B200CholeskyFactorization

"""
    B200CholeskyFactorization(; ordering = nothing, device = 0)

Sparse Cholesky on one B200. A new `LinearCache` analyses the pattern once; `cache.A = Q_new` with the same pattern
(the Newton loop's `_update_linsolve_cache!`) refactorizes numerically only.
"""
struct B200CholeskyFactorization{O} <: LinearSolve.AbstractSparseFactorization
    ordering::O
    device::Int
end
B200CholeskyFactorization(; ordering = nothing, device::Integer = 0) = B200CholeskyFactorization(ordering, Int(device))

_csc(A::SparseMatrixCSC{Float64, Int}) = A
_csc(A::Symmetric) = SparseMatrixCSC{Float64, Int}(sparse(A))
_csc(A::AbstractMatrix) = SparseMatrixCSC{Float64, Int}(sparse(A))

# LinearSolve hooks (third-party API): build the factorization lazily at the first solve!, like CHOLMODFactorization
LinearSolve.init_cacheval(::B200CholeskyFactorization, A, b, u, Pl, Pr, maxiters::Int, abstol, reltol, verbose, assumptions) = nothing

function SciMLBase.solve!(cache::LinearSolve.LinearCache, alg::B200CholeskyFactorization; kwargs...)
    A = cache.A
    if cache.isfresh
        be = LinearSolve.@get_cacheval(cache, :B200CholeskyFactorization)
        Q = _csc(A)
        if be === nothing || be.nnz != nnz(Q)
            be = B200Backend(Q; ordering = alg.ordering, device = alg.device)     # symbolic + numeric
        else
            refactorize!(be, Symmetric(Q))                                        # same pattern: numeric only
        end
        cache.cacheval = be
        cache.isfresh = false
    end
    be = LinearSolve.@get_cacheval(cache, :B200CholeskyFactorization)
    cache.u .= backend_solve(be, cache.b)
    return SciMLBase.build_linear_solution(alg, cache.u, nothing, cache)
end

# ---- the package's dispatch tables ------------------------------------------------------------------------------------
GaussianMarkovRandomFields.supports_selinv(::B200CholeskyFactorization) = Val{true}()                 # solvers/selinv.jl:16-29
GaussianMarkovRandomFields.supports_backward_solve(::B200CholeskyFactorization) = Val{true}()         # solvers/backward_solve.jl:14-27

_backend(linsolve) = LinearSolve.@get_cacheval(linsolve, :B200CholeskyFactorization)

GaussianMarkovRandomFields._selinv_diag_impl(linsolve, ::B200CholeskyFactorization) = get_selinv_diag(_backend(linsolve))   # selinv.jl:70-73
GaussianMarkovRandomFields._selinv_impl(linsolve, ::B200CholeskyFactorization) = Symmetric(get_selinv(_backend(linsolve))) # selinv.jl:107-110
GaussianMarkovRandomFields._backward_solve_impl(linsolve, x, ::B200CholeskyFactorization) =                                # backward_solve.jl:50-53
    backend_backward_solve(_backend(linsolve), x isa VecOrMat ? x : collect(x))
GaussianMarkovRandomFields._logdet_cov_impl(linsolve, ::B200CholeskyFactorization) = -compute_logdet(_backend(linsolve))   # logdet.jl:27-31

# the library wants the full symmetric CSC (both triangles), like GMRFWorkspace.Q (gmrf_workspace.jl:32)
GaussianMarkovRandomFields.prepare_for_linsolve(A::AbstractMatrix, ::B200CholeskyFactorization) = _csc(A)                  # linsolve_utils.jl:10-24
GaussianMarkovRandomFields.algorithm_applicable(::B200CholeskyFactorization, ::SparseMatrixCSC) = Val{true}()
GaussianMarkovRandomFields.algorithm_applicable(::B200CholeskyFactorization, ::AbstractMatrix) = Val{false}()             # linsolve_utils.jl:43-56

# Newton loop: same pattern, new values -> numeric refactorization at the next solve! (condition/gaussian_approximation.jl:61-76)
function GaussianMarkovRandomFields._update_linsolve_cache_inner!(cache, Q, ::B200CholeskyFactorization)
    return cache.A = _csc(Q)           # LinearSolve's setproperty! marks the cache fresh
end

end # module
