"""The FP64 sequences the kernels use instead of library calls (branch-free sqrt, shared-reciprocal
quotients with Markstein corrections) must reproduce the IEEE operation bit for bit: that is what keeps
the L0Box / L1Box iprox! kernels bit-exact against the oracle while issuing a third of the divisions."""
import ctypes as C

import pytest

from gpu_util import DEV, sp
from shiftedprox import _lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 20261018])
def test_math_building_blocks_are_ieee_exact(seed):
    out = (C.c_int64 * 3)()
    L.call("spx_selftest_math", sp.context(DEV), C.c_int64(1 << 26), C.c_uint64(seed), out)
    assert list(out) == [0, 0, 0], list(out)
