"""Known-answer tests for the Philox port and the PyTorch CUDA stream mapping (CPU only)."""
import ctypes
import os

import numpy as np
import torch

from oracle import philox as ph


KAT = [  # Random123 kat_vectors: philox4x32-10
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, exp in KAT:
        out = ph.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert tuple(int(v) for v in out) == exp


def test_philox_matches_c_oracle():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libvq_oracle.so")
    if not os.path.exists(path):
        import pytest
        pytest.skip("C oracle not built")
    lib = ctypes.CDLL(path)
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2**32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2**32, size=(64, 2), dtype=np.uint64).astype(np.uint32)
    got = ph.philox4x32_10(ctr, key)
    for i in range(64):
        out = (ctypes.c_uint32 * 4)()
        lib.philox4x32_10((ctypes.c_uint32 * 4)(*ctr[i].tolist()), (ctypes.c_uint32 * 2)(*key[i].tolist()), out)
        assert list(out) == got[i].tolist()


def test_launch_policy_b200():
    # 148 SMs x 2048 threads: full grid is 1184 blocks; one exponential_ over [65536, 8192]
    # consumes ((numel-1)//(256*1184*4)+1)*4 offsets.
    assert ph.launch_grid(65536 * 8192, 148) == 1184
    assert ph.launch_grid(65536, 148) == 256
    assert ph.offset_increment(65536, 148) == 4
    assert ph.offset_increment(65536 * 8192, 148) == ((65536 * 8192 - 1) // (256 * 1184 * 4) + 1) * 4
    # the torch mapping has a period of exactly 148 rows of 8192: element (row, col) uses
    # thread (row % 37) * 8192 + col, call row // 148, lane (row % 148) // 37.
    assert 256 * 1184 == 37 * 8192


def test_uniform_ranges():
    u = ph.u32_to_uniform(np.array([0, 1, 2**31, 2**32 - 1], dtype=np.uint32))
    assert u[0] > 0 and u[-1] == np.float32(1.0)
    r = ph.rand(1000, seed=42, offset=0, sm_count=148)
    assert r.min() >= 0 and r.max() < 1
    ri = ph.randint(1000, 8192, seed=42, offset=0, sm_count=148)
    assert ri.min() >= 0 and ri.max() < 8192


def test_multinomial_is_argmax_p_over_q_cpu():
    """torch.multinomial(p,1) == argmax(p / Exp(1)) with the generator consumed identically
    (the structure the CUDA kernel relies on; re-verified against torch CUDA in the gpu tests)."""
    g = torch.Generator().manual_seed(3)
    p = torch.softmax(torch.randn(64, 100, generator=g) * 3, dim=-1)
    torch.manual_seed(9)
    a = torch.multinomial(p, 1)[:, 0]
    st = torch.get_rng_state()
    torch.manual_seed(9)
    q = torch.empty_like(p).exponential_(1)
    b = torch.argmax(p / q, dim=-1)
    assert torch.equal(a, b)
    assert torch.equal(st, torch.get_rng_state())
