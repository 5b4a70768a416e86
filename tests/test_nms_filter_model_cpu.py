"""float32 model of the experimental NMS division filter (csrc/nms.cu: suppresses<DIM, FILTER=true>): deciding `inter / uni > thresh` by the
band test must agree with the correctly rounded division for every input, including inter within a few ulps of thresh * uni."""
import numpy as np
import pytest

f = np.float32


def _decisions(inter, uni, t):
    inter, uni, t = inter.astype(f), uni.astype(f), f(t)
    exact = (inter / uni) > t
    p = (t * uni).astype(f)
    hi, lo = (p * f(1.000001)).astype(f), (p * f(0.999999)).astype(f)
    in_range = (t >= f(1e-30)) & (t <= f(1)) & (uni >= f(1)) & (uni <= f(1e30))
    filt = np.where(inter > hi, True, np.where(inter < lo, False, exact))
    return np.where(in_range, filt, exact), exact


@pytest.mark.parametrize("t", [0.5, 1e-5, 0.7, 1.0 / 3.0, 1.0, 1e-30, 0.999999])
def test_band_filter_equals_ieee_division(t):
    rs = np.random.RandomState(int(t * 1000) % 97)
    uni = np.exp(rs.uniform(0, 40, size=300000)).astype(f)
    got, want = _decisions((uni * rs.uniform(0, 1.2, size=uni.size)).astype(f), uni, t)
    assert np.array_equal(got, want)
    inter = (f(t) * uni).astype(f)            # adversarial: walk up to 40 ulps away from thresh * uni in both directions
    k = rs.randint(-40, 41, size=uni.size)
    for _ in range(40):
        inter = np.where(k > 0, np.nextafter(inter, f(np.inf)), np.where(k < 0, np.nextafter(inter, f(0)), inter))
        k = k - np.sign(k)
    got, want = _decisions(inter, uni, t)
    assert np.array_equal(got, want)
