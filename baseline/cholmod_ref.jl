# cholmod_ref.jl -- times the REFERENCE's own CHOLMOD path (GMRFWorkspace / CHOLMODBackend,
# src/workspace/gmrf_workspace.jl, src/workspace/backend.jl) on the matrices exported by baseline/export_configs.py.
# This image has no Julia, so the script has never been run here; it is the recipe BASELINE.md section 3 (item 3) asks
# for, to be executed on any machine with Julia >= 1.10 and GaussianMarkovRandomFields.jl installed:
#
#     julia -t auto baseline/cholmod_ref.jl baseline/_matrices/config4.bin [more.bin ...]
#
# One JSON line per matrix: seconds for symbolic + first numeric, numeric refactorization + logdet (the bench metric),
# one solve, one half solve, selinv diagonal; once with the permutation the B200 arm uses (ordering = perm) and once with
# CHOLMOD's default ordering. BLAS/CHOLMOD thread counts are printed with the result.
using GaussianMarkovRandomFields
using GaussianMarkovRandomFields: workspace_solve, backward_solve, selinv_diag   # not exported (test/workspace/test_gmrf_workspace.jl:2)
using SparseArrays, LinearAlgebra, Printf

function read_matrix(path)
    open(path, "r") do io
        n, nz, has_perm = read(io, Int64), read(io, Int64), read(io, Int64)
        colptr = Vector{Int64}(undef, n + 1); read!(io, colptr)
        rowval = Vector{Int64}(undef, nz);    read!(io, rowval)
        nzval  = Vector{Float64}(undef, nz);  read!(io, nzval)
        perm = has_perm == 1 ? (p = Vector{Int64}(undef, n); read!(io, p); p) : nothing
        return SparseMatrixCSC{Float64, Int}(n, n, colptr, rowval, nzval), perm
    end
end

best(f, reps) = minimum(begin t = time_ns(); f(); (time_ns() - t) / 1e9 end for _ in 1:reps)

function time_one(Q, ordering, label; reps = 3)
    n = size(Q, 1)
    t0 = time_ns()
    ws = ordering === nothing ? GMRFWorkspace(Q) : GMRFWorkspace(Q; ordering = ordering)
    t_setup = (time_ns() - t0) / 1e9
    nz = copy(nonzeros(Q))
    scale = Ref(1.0)
    refactor() = (scale[] *= 1.0000001; update_precision_values!(ws, nz .* scale[]); logdet(ws))
    refactor()                                                       # warm-up
    t_factor = best(refactor, reps)
    b = randn(n)
    workspace_solve(ws, b); t_solve = best(() -> workspace_solve(ws, b), reps)
    backward_solve(ws, b);  t_half = best(() -> backward_solve(ws, b), reps)
    t_selinv = best(() -> (update_precision_values!(ws, nz); selinv_diag(ws)), 1)
    @printf("{\"ordering\": \"%s\", \"n\": %d, \"nnz_q\": %d, \"setup_s\": %.3f, \"refactorize_logdet_s\": %.4f, \"solve_s\": %.4f, \"half_solve_s\": %.4f, \"selinv_diag_s\": %.3f, \"logdet\": %.12g, \"julia_threads\": %d, \"blas_threads\": %d}\n",
            label, n, nnz(Q), t_setup, t_factor, t_solve, t_half, t_selinv, logdet(ws), Threads.nthreads(), BLAS.get_num_threads())
end

for path in ARGS
    Q, perm = read_matrix(path)
    println("# ", path)
    perm === nothing || time_one(Q, perm, "perm of the B200 arm")
    time_one(Q, nothing, "CHOLMOD default")
end
