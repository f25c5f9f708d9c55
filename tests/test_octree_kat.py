"""Known-answer vectors of attcs/Octree @23bda72 (the third-party library whose RayIntersectedAll defines the
reference's hit set and order) replayed against the oracle's restatement of IsRayHit + sort-by-(distance,id)
(orc_ray_hits_generic).  Vectors restated from external/Octree/unittests/general.tests.cpp:839-1063 and
unittests/adaptor.tests.cpp:255-256 (inputs and expected ids only)."""
import numpy as np

from libs import dp, u64p

DIAG4 = [[0, 0, 1, 1], [1, 1, 2, 2], [2, 2, 3, 3], [3, 3, 4, 4]]       # four unit boxes on the diagonal
FIVE2 = [[2, 0, 3, 1], [3, 1, 4, 2], [3, 2, 4, 3], [2, 3, 4, 4], [2.5, 2.5, 3.5, 3.5]]
FIVE3 = [[2, 0, 1, 3, 1, 2], [3, 1, 1, 4, 2, 2], [3, 2, 1, 4, 3, 2], [2, 3, 1.5, 4, 4, 2], [2.5, 2.5, 1.5, 3.5, 3.5, 2]]


def hits(orc, dim, boxes, o, h):
    b = np.ascontiguousarray(boxes, np.float64)
    ids = np.zeros(16, np.uint64)
    n = orc.orc_ray_hits_generic(dim, dp(b), b.shape[0], dp(np.array(o, np.float64)), dp(np.array(h, np.float64)), u64p(ids), 16)
    return [int(i) for i in ids[:n]]


def test_ray_intersected_all_kats(orc):
    assert hits(orc, 2, FIVE2, [2.7, 2.6], [1.0, 1.0]) == [4, 2, 3]            # RayIntersectedAll_2D_General_234 (ordered)
    assert sorted(hits(orc, 2, FIVE2, [2.6, 2.8], [1.0, 1.0])) == [3, 4]       # RayIntersectedAll_2D_General_34 (permutation)
    assert sorted(hits(orc, 3, FIVE3, [2.6, 2.8, 1.0], [1.0, 1.0, 1.0])) == [3, 4]   # RayIntersectedAll_3D_General_34


def test_ray_intersected_first_kats(orc):
    """RayIntersectedFirst == first element of the (distance, id)-sorted list; None == empty list"""
    first = lambda o, h: (hits(orc, 2, DIAG4, o, h) or [None])[0]
    assert first([0.0, -1.0], [1.0, 0.0]) is None       # XRay None
    assert first([5.0, 0.0], [0.0, 1.0]) is None        # YRay None
    assert first([4.2, 4.4], [1.0, 1.0]) is None        # YRay2 None
    assert first([1.5, 2.6], [-1.0, -1.0]) is None      # InsideTheTree_Neg None
    assert first([7.0, 2.0], [-1.0, -1.0]) is None      # OutsideTheTree_Neg None
    assert first([0.0, 0.5], [1.0, 0.0]) == 0           # XRay 0
    assert first([0.0, 2.0], [1.0, 0.0]) == 1           # XRay 1 (grazes the shared edge y = 2: closed intervals)
    assert first([1.5, 1.0], [1.0, 0.0]) == 1           # XRayInsideTheTreePos 1
    assert first([1.5, 1.0], [-1.0, 0.0]) == 1          # XRayInsideTheTreeNeg 1
    assert first([3.5, 3.5], [-0.5, -1.0]) == 3         # XRayInsideTheTreeNeg 3
