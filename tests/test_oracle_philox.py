"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) and eps-stream sanity."""
import numpy as np

from oracle import philox


def _kat(ctr, key):
    out = philox.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
    return [int(v) for v in out]


def test_random123_known_answers():
    assert _kat([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _kat([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _kat([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normal_stream_moments_and_independence():
    a = philox.normals(seed=1234, step=0, n_elems=200_000)
    b = philox.normals(seed=1234, step=1, n_elems=200_000)
    c = philox.normals(seed=1235, step=0, n_elems=200_000)
    assert a.dtype == np.float32 and np.isfinite(a).all()
    for v in (a, b, c):
        assert abs(v.mean()) < 0.01 and abs(v.std() - 1) < 0.01
        assert abs(((v - v.mean()) ** 3).mean()) < 0.03          # skew
        assert abs((v ** 4).mean() - 3) < 0.1                    # kurtosis
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.01 and abs(np.corrcoef(a, c)[0, 1]) < 0.01
    # prefix property: a shorter request is a prefix of a longer one
    assert np.array_equal(philox.normals(1234, 0, 10), a[:10])
