"""CPU: numpy restatements of the helper surface (oracle/oracle.py) against the golden vectors the
unmodified reference produced (tests/golden/tools.npz, oracle/make_golden.py:gen_tools)."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle.make_golden import tools_inputs


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "tools.npz")


@pytest.fixture(scope="module")
def inp():
    return tools_inputs()


def test_filter_data(gold, inp):
    for d in ("up", "down"):
        assert np.array_equal(orc.filter_data(inp["fd_x"], d), gold[f"fd_{d}"])


def test_detect_onset_region(gold, inp):
    a = [orc.detect_onset_region(x, int(o)) for x, o in zip(inp["or_x"], inp["or_on"])]
    b = [orc.detect_onset_region(x, int(o), 128, 7, 0.3) for x, o in zip(inp["or_x"], inp["or_on"])]
    assert a == gold["or_a"].tolist() and b == gold["or_b"].tolist()


def test_lfilter_orders(gold, inp):
    from scipy import signal as sig

    for tag, (cut, order, bt) in {"lo2": (3000, 2, "low"), "hi3": (500, 3, "high")}.items():
        b, a = sig.butter(order, cut, btype=bt, output="ba", fs=96000)
        zi = np.zeros((order, 4), np.float32)
        for k, blk in enumerate(inp["bw_x"]):
            y, zi = orc.lfilter_f32(np.float32(b), np.float32(a), blk, zi)
            assert np.array_equal(y, gold[f"bw_{tag}"][k])


def test_find_lag(gold, inp):
    got = [orc.find_lag(a, b) for a, b in zip(inp["fl_a"][:10], inp["fl_b"][:10])]
    assert got == gold["fl"][:10].tolist()
