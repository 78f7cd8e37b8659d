# run_reference_tests_on_b200.jl -- the reference's OWN workspace test files as the parity suite of the B200 backend.
#
# The files under test/workspace/ build their workspaces with `GMRFWorkspace(Q; kw...)`, which is hard-wired to CHOLMOD
# (src/workspace/gmrf_workspace.jl:66-84). This harness re-points that one constructor (and the pool's) at `B200Backend`
# and then includes the reference's test files unchanged, so every assertion they make -- dense-LinearAlgebra identities
# at the reference's own tolerances -- is made against libgmrf_b200.so. Not executed in this repository (no Julia in the
# image); the Python mirrors under tests/ assert the same things through the same C-ABI.
#
#     GMRF_B200_LIB=/path/to/libgmrf_b200.so julia --project=/path/to/GaussianMarkovRandomFields.jl \
#         run_reference_tests_on_b200.jl /path/to/GaussianMarkovRandomFields.jl
using Test
using GaussianMarkovRandomFields
using SparseArrays, LinearAlgebra

include(joinpath(@__DIR__, "B200Backend.jl"))
using .GaussianMarkovRandomFieldsB200

const GMRFs = GaussianMarkovRandomFields
const REF = length(ARGS) >= 1 ? ARGS[1] : pkgdir(GMRFs)

# The default constructor now builds a B200 workspace; `ordering = ...` keeps working (resolved on the host by the package's
# own `ordering_permutation`, handed over as a permutation).
function GMRFs.GMRFWorkspace(Q::SparseMatrixCSC{T}; backend_kwargs...) where {T}
    return GMRFs.GMRFWorkspace(Q, B200Backend; backend_kwargs...)
end

@testset "reference workspace tests on B200Backend" begin
    for f in ("test_gmrf_workspace.jl", "test_backend_ordering.jl", "test_workspace_gmrf.jl", "test_workspace_constrained.jl",
              "test_workspace_gaussian_approximation.jl", "test_workspace_latent_models.jl", "test_workspace_pool.jl",
              "test_precision_logdet.jl", "test_workspace_autodiff.jl")
        @testset "$f" begin
            include(joinpath(REF, "test", "workspace", f))
        end
    end
end
# Known differences to expect: assertions that peek at CHOLMOD-specific fields (`ws.backend.factor`,
# test_gmrf_workspace.jl:214-219 reads `selinv_cache` / `selinv_diag_cache`, which B200Backend mirrors by name) and
# test_cliquetrees_backend.jl, which constructs its backend explicitly and is therefore not re-pointed.
