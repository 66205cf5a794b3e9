"""Dry run of the estimator-level GPU tests on the CPU stand-ins (tests/fake_ops.py).

The bodies of the `-m gpu` tests that drive the public estimators are executed here with `torchdr_b200.ops` replaced
by oracle-backed stand-ins and the device set to the CPU.  What this checks is the HOST code those tests go through
(input handling, duplicate processing, hooks, loops, optimiser / scheduler plumbing, error messages) and the test bodies
themselves — so that a host-side change cannot break the GPU suite unnoticed between two GPU sessions.  Numerical
parity of the CUDA kernels is what the real `-m gpu` run checks.
"""

import pytest

import fake_ops
import test_gpu_parity as gpu_tests

ESTIMATOR_TESTS = [
    "test_estimators_end_to_end",
    "test_umap_estimator_parity_hooks",
    "test_estimator_edge_cases",
    "test_discard_nns_estimators_on_gpu",
    "test_generic_optimizers_on_gpu",
]


@pytest.mark.parametrize("name", ESTIMATOR_TESTS)
def test_gpu_test_body_runs_on_cpu_stand_ins(name, monkeypatch):
    fake_ops.install(monkeypatch)
    fake_ops.install_entropic(monkeypatch)
    monkeypatch.setattr(gpu_tests, "DEV", "cpu")
    monkeypatch.setattr(gpu_tests, "_cuda", lambda x: x)
    getattr(gpu_tests, name)()
