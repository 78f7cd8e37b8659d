"""gmrf_b200: host-side mirror of GaussianMarkovRandomFields.jl's workspace/solver API over libgmrf_b200.so."""
