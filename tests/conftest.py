"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on the CPU-only build container (oracle vs goldens, host logic, C-ABI symbol
checks, gloo world_size-2 sharding); `-m gpu` runs the parity tests proper through the C-ABI on a
B200.  Nothing here reads /root/reference: the goldens live in tests/golden/.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _split_cols(g, prefix):
    v, l, nr = g[prefix + "_values"], g[prefix + "_lengths"], g[prefix + "_nruns"]
    off = np.concatenate(([0], np.cumsum(nr)))
    return [(v[off[i]:off[i + 1]], l[off[i]:off[i + 1]]) for i in range(len(nr))]


class Golden:
    """The reference's golden objects (tests/golden/reference_goldens.npz) in convenient form."""

    def __init__(self):
        from oracle import derfinder_oracle as O
        self.g = np.load(os.path.join(ROOT, "tests", "golden", "reference_goldens.npz"))
        g = self.g
        self.cov_rle = _split_cols(g, "genomeData_cov")
        self.raw_rle = _split_cols(g, "genomeDataRaw_cov")
        self.cov = np.stack([O.rle_decode(*c) for c in self.cov_rle], axis=1)  # [1434 x 31] int32
        self.position = (g["genomeData_pos_values"], g["genomeData_pos_lengths"])
        self.pop = g["genomeInfo_pop"]
        self.gender = g["genomeInfo_gender"]
        self.mod = g["chr21_mod"]
        self.mod0 = g["chr21_mod0"]
        self.fstats = O.rle_decode(g["genomeFstats_values"], g["genomeFstats_lengths"])
        self.fstats_matrix = O.rle_decode(g["chr21_fstats_values"], g["chr21_fstats_lengths"])

    def __getitem__(self, k):
        return self.g[k]


@pytest.fixture(scope="session")
def golden():
    return Golden()
