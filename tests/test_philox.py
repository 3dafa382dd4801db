"""Known-answer test of the numpy Philox4x32-10 restatement (oracle/philox.py) against the
Random123 `kat_vectors` entries for philox4x32 with 10 rounds.  The CUDA generator is then
checked against this restatement bit for bit in tests/test_gpu_parity.py."""
import numpy as np

from oracle.philox import philox4x32_10

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_random123_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10(ctr, key)
        assert [int(x) for x in got] == want


def test_vectorised_matches_scalar():
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2 ** 32, size=(50, 4), dtype=np.uint64)
    key = rng.integers(0, 2 ** 32, size=(50, 2), dtype=np.uint64)
    batch = philox4x32_10(ctr, key)
    for i in range(50):
        assert np.array_equal(batch[i], philox4x32_10(ctr[i], key[i]))
