"""Exhaustive check of the device expf (gtb_dev.cuh: expf_glibc, a restatement of glibc 2.39's FMA-variant expf) against the
host libm over ALL 2^32 float bit patterns.  FP16 greedy identity hangs on every bit of it (hard part 4: SiLU
and the softmax call expf, gten/ops.h:692, 985), and it holds only where the box's glibc computes what 2.39 computes --
this test is that check, on the box."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def test_expf_all_2_32_inputs():
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    port = oracle.port()
    chunk = 1 << 24
    bad = 0
    first_bad = None
    for c in range(256):
        first = c * chunk
        got = capi.selftest_expf(first, chunk)
        want = port.expf_bits_range(first, chunk)
        gb, wb = got.view(np.uint32), want.view(np.uint32)
        diff = gb != wb
        if diff.any():
            # NaN inputs: any NaN result matches any NaN result (payloads are not part of the contract: no consumer reads them)
            diff &= ~(np.isnan(got) & np.isnan(want))
        n = int(diff.sum())
        if n and first_bad is None:
            i = int(np.argmax(diff))
            first_bad = (hex(first + i), hex(int(gb[i])), hex(int(wb[i])))
        bad += n
    assert bad == 0, f"{bad} of 2^32 inputs differ from the host libm; first: input bits {first_bad[0]}, device {first_bad[1]}, host {first_bad[2]}"
