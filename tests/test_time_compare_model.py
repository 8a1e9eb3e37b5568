"""CPU: the two arithmetic identities the thread-per-lookahead kernel's tick loop rests on (ddls_b200/csrc/ramp_lookahead_thread.cuh),
on IEEE doubles including subnormals, equal values, zero and +inf:

  * remaining times are non-negative, so comparing them is comparing their bit patterns as unsigned 64-bit integers;
  * the reference's "x -= min(tick, x); if x == 0: completed" (JOB:555-556, 561-562) is "x <= tick": for x > tick >= 0 the
    difference x - tick is never rounded to zero (gradual underflow), and for x <= tick it is exactly +0."""
import numpy as np
from hypothesis import given, settings, strategies as st

nonneg = st.one_of(st.floats(min_value=0.0, allow_nan=False, allow_infinity=True), st.sampled_from([0.0, 5e-324, 2.2250738585072014e-308, 1.0, np.inf]),
                   st.floats(min_value=0.0, max_value=1e-300, allow_nan=False))


def bits(x):
    return int(np.float64(x).view(np.uint64))


@settings(max_examples=3000, deadline=None)
@given(nonneg, nonneg)
def test_unsigned_bit_patterns_order_like_the_values(a, b):
    assert (a < b) == (bits(a) < bits(b))
    assert (a == b) == (bits(a) == bits(b)) or (a == 0.0 and b == 0.0)       # only +0.0 occurs (costs are canonicalised with + 0.0)
    assert min(a, b) == (a if bits(a) < bits(b) else b)


@settings(max_examples=3000, deadline=None)
@given(nonneg, nonneg)
def test_completion_test_is_a_comparison(x, tick):
    if np.isinf(x) and np.isinf(tick):
        return                                        # inf - inf is NaN in the reference too; an infinite tick is an error (RCE:462)
    m = x if x < tick else tick                       # Python's min(tick, x) for these operands
    left = np.float64(x) - np.float64(m)
    assert (left == 0.0) == (x <= tick)
    if x > tick:
        assert left > 0.0 and left == np.float64(x) - np.float64(tick)


def test_close_neighbours_do_not_collapse():
    for base in (5e-324, 1e-310, 2.2250738585072014e-308, 1.0, 1e300):
        x = np.nextafter(np.float64(base), np.inf)
        assert x > base and x - np.float64(base) > 0.0
