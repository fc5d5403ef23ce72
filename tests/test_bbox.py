"""bbox_overlaps_cython (cython/bbox.pyx:15-55): oracle vs golden vectors generated from the
compiled reference (CPU), CUDA kernel vs oracle and golden (GPU)."""
import os

import numpy as np
import pytest

from tests.golden import make_bbox_golden as gen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bbox_golden.npz")


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_oracle_matches_reference_golden(oracle, golden):
    small_b, small_q, big_b, big_q = gen.cases()
    assert np.array_equal(_bits(small_b), _bits(golden["small_boxes"])), "numpy generator drift"
    got = oracle.bbox_overlaps(golden["small_boxes"], golden["small_query"])
    assert np.array_equal(_bits(got), _bits(golden["small_overlaps"]))
    assert got[6, 5] == 1.0 and (got[1] == 0).all() and (got[2] == 0).all()
    assert gen.digest(big_b, big_q) == str(golden["big_inputs_digest"])
    assert gen.digest(oracle.bbox_overlaps(big_b, big_q)) == str(golden["big_overlaps_digest"])


def test_oracle_equals_reference_build(oracle):
    from oracle import ref
    if not ref.bbox_available():
        pytest.skip("oracle/_ref/bbox_ref*.so not built (no /root/reference on this box)")
    for seed, n, k in ((1, 1, 1), (2, 130, 257), (3, 64, 1)):
        b, q = gen.boxes(seed, n, n >= 8), gen.boxes(seed + 100, k, k >= 8)
        assert np.array_equal(_bits(oracle.bbox_overlaps(b, q)), _bits(ref.bbox_overlaps_cython(b, q)))


@pytest.mark.gpu
@pytest.mark.parametrize("n,k", [(1, 1), (37, 21), (16, 128), (17, 129), (3000, 700), (5, 1000)])
def test_cuda_matches_oracle(oracle, cuda, n, k):
    from dspnet_b200.bbox import bbox_overlaps_cython
    b, q = gen.boxes(900 + n, n, n >= 8), gen.boxes(1900 + k, k, k >= 8)
    want = oracle.bbox_overlaps(b, q)
    got = bbox_overlaps_cython(b, q)
    assert got.dtype == np.float64 and got.shape == (n, k)
    assert np.array_equal(_bits(got), _bits(want))
    import torch
    dev = bbox_overlaps_cython(torch.from_numpy(b).to(cuda), torch.from_numpy(q).to(cuda))
    assert dev.is_cuda and np.array_equal(_bits(dev.cpu().numpy()), _bits(want))


@pytest.mark.gpu
def test_cuda_matches_reference_golden(cuda, golden):
    from dspnet_b200.bbox import bbox_overlaps_cython
    got = bbox_overlaps_cython(golden["small_boxes"], golden["small_query"])
    assert np.array_equal(_bits(got), _bits(golden["small_overlaps"]))
    _, _, big_b, big_q = gen.cases()
    assert gen.digest(bbox_overlaps_cython(big_b, big_q)) == str(golden["big_overlaps_digest"])
    assert bbox_overlaps_cython(np.zeros((0, 4)), big_q).shape == (0, 700)
