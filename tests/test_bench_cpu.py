"""bench.py host-side pieces that do not need a GPU: the roofline object of the JSON line."""
import json

import bench


def test_roofline_object_has_the_contract_keys():
    tag = (0, (2, 36, 128, 128, 128), (36, 36, 3, 3, 3), (1, 1, 1), 0)
    dom = (2.0 * 2 * 128 ** 3 * 36 * 36 * 27, 2.5, tag, 1.0)
    btag = (3, (2, 36, 128, 128, 128), (36, 36, 3, 3, 3), (1, 1, 1), 2)
    slow = (2 * dom[0], 6.3, btag, 1.0)
    traffic = {"input_shape": [2, 36, 128, 128, 128], "dram_bytes": 1.378e9, "source": "profiles/r02_ncu_conv_tcw_p0_36.txt"}
    r = bench.make_roofline(3.6e12, 52.0, 229, 66.7, dom, 1444.3, True, slow, traffic)
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "slowest_launch"):
        assert k in r
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["traffic"] == 1.378e9 and "r02" in r["traffic_source"]
    assert r["slowest_launch"]["ms"] == 6.3 and "fused backward" in r["slowest_launch"]["kernel"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] < 1
    assert abs(r["achieved"] - dom[0] / 2.5 / 1e9) < 1e-9
    agg = r["all_conv_launches"]
    assert abs(agg["achieved"] - 3.6e12 / 52e-3 / 1e12) < 1e-9 and abs(agg["conv_share_of_step"] - 52.0 / 66.7) < 1e-12
    json.dumps(r)
    # no per-launch tag (e.g. SIMT-only run): the aggregate is reported at the top level; no conv time at all: no roofline
    r2 = bench.make_roofline(3.6e12, 52.0, 229, 66.7, None, 1400.0, False)
    assert r2["traffic"] is None and "fallback" in r2["peak_source"] and r2["frac"] == r2["all_conv_launches"]["frac"]
    assert bench.make_roofline(3.6e12, None, 0, 66.7, None, 1400.0, False) is None
