"""Developer tool (NOT collected by pytest, never part of a parity claim): runs the `[b200]` arms of the two-arm host-logic
tests on a machine WITHOUT a GPU by swapping `B200Backend` for a dense stand-in that has B200Backend's exact constructor
signature. It proves nothing about the CUDA path -- it only catches keyword / plumbing mistakes in GPU arms before they
cost GPU minutes. Usage: python tests/dry_run_gpu_arms.py"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gaussianmarkovrandomfields.jl_b200"), os.path.join(ROOT, "tests")]

import gmrf_b200.backend as backend  # noqa: E402
from dense_backend import DenseBackend  # noqa: E402


class StrictSignatureDense(DenseBackend):
    def __init__(self, Q, ordering=None, device: int = 0, check: bool = False, factorize: bool = True):
        super().__init__(Q)

    def permutation(self):
        return np.arange(self.n)

    def pin_host_buffer(self, arr):
        return True

    def close(self):
        pass


if __name__ == "__main__":
    backend.B200Backend = StrictSignatureDense
    files = ["test_zzz_latent_model_integration.py", "test_zz_selinv_traces.py", "test_workspace_gmrf.py", "test_gmrf_boundary_b.py"]
    sys.exit(pytest.main(["-q", "-m", "gpu", "-p", "no:cacheprovider", "-k", "b200"] + [os.path.join(ROOT, "tests", f) for f in files]))
