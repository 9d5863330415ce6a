"""CPU oracle for the random streams the sampler must reproduce.

TEST INFRASTRUCTURE ONLY (see oracle/paella_oracle.py header).

The reference draws its randomness through PyTorch's CUDA generator
(``torch.randint`` ref/src/utils.py:37, ``torch.multinomial`` :50,
``torch.rand_like`` ref/src/modules.py:279).  The arithmetic is PyTorch's, not
the reference's; the installed headers pin it:
  torch/include/ATen/native/cuda/DistributionTemplates.h:33-38,50-62  launch policy + philox offset
  .../DistributionTemplates.h:65-89                                   grid-stride element <-> (thread, call, lane) map
  .../DistributionTemplates.h:282-316                                 randint -> curand4, ``val % range + base``
  .../DistributionTemplates.h:485-506                                 uniform: ``rand*range+from``, ``== to -> from``
  torch/include/ATen/core/TransformationHelper.h:129-146              exponential (CUDA branch)
  /usr/local/cuda/include/curand_philox4x32_x.h                       Philox4x32-10, counter/key layout
  /usr/local/cuda/include/curand_kernel.h (_curand_uniform)           u = x*2^-32 + 2^-33

Philox is integer arithmetic, so this numpy port is exact; it is checked
against the Random123 known-answer vectors in tests/test_oracle_philox.py.
"""
from __future__ import annotations

import numpy as np

PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)

BLOCK = 256          # DistributionTemplates.h:33
UNROLL = 4           # float / uint32 outputs: one curand*4 call feeds 4 elements


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [...,4] uint32, key [...,2] uint32 -> [...,4] uint32."""
    c = [ctr[..., i].astype(np.uint32) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32)
    k1 = key[..., 1].astype(np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c[0].astype(np.uint64) * PHILOX_M0
            p1 = c[2].astype(np.uint64) * PHILOX_M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + PHILOX_W0).astype(np.uint32)
            k1 = (k1 + PHILOX_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def launch_grid(numel: int, sm_count: int, max_threads_per_sm: int = 2048) -> int:
    """calc_execution_policy: grid.x (DistributionTemplates.h:50-58)."""
    return min(sm_count * (max_threads_per_sm // BLOCK), (numel + BLOCK - 1) // BLOCK)


def offset_increment(numel: int, sm_count: int) -> int:
    """Philox offset consumed by one distribution kernel over ``numel`` elements
    (DistributionTemplates.h:59-60)."""
    grid = launch_grid(numel, sm_count)
    return ((numel - 1) // (BLOCK * grid * UNROLL) + 1) * 4


def raw_u32(numel: int, seed: int, offset: int, sm_count: int) -> np.ndarray:
    """The uint32 that PyTorch's grid-stride kernel hands to element e, e = 0..numel-1."""
    grid = launch_grid(numel, sm_count)
    stride = BLOCK * grid
    e = np.arange(numel, dtype=np.int64)
    tid = e % stride
    j = e // stride
    call, lane = j // UNROLL, j % UNROLL
    ctr_lo = np.uint64(offset // 4) + call.astype(np.uint64)
    ctr = np.stack([(ctr_lo & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                    (ctr_lo >> np.uint64(32)).astype(np.uint32),
                    (tid & 0xFFFFFFFF).astype(np.uint32),
                    (tid >> 32).astype(np.uint32)], axis=-1)
    key = np.empty((numel, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    out = philox4x32_10(ctr, key)
    return out[np.arange(numel), lane]


def u32_to_uniform(x: np.ndarray) -> np.ndarray:
    """_curand_uniform: (0,1] fp32."""
    return (x.astype(np.float32) * np.float32(2.3283064365386963e-10)
            + np.float32(2.3283064365386963e-10 / 2.0)).astype(np.float32)


def randint(numel: int, high: int, seed: int, offset: int, sm_count: int) -> np.ndarray:
    """torch.randint(0, high, ...) on CUDA for high < 2^28 (32-bit path)."""
    return (raw_u32(numel, seed, offset, sm_count).astype(np.uint64) % np.uint64(high)).astype(np.int64)


def rand(numel: int, seed: int, offset: int, sm_count: int) -> np.ndarray:
    """torch.rand / rand_like fp32 on CUDA: [0,1)."""
    u = u32_to_uniform(raw_u32(numel, seed, offset, sm_count))
    return np.where(u == np.float32(1.0), np.float32(0.0), u)


def exponential(numel: int, seed: int, offset: int, sm_count: int) -> np.ndarray:
    """Tensor.exponential_(1) fp32 on CUDA.  NB: the device uses the fast ``__logf``
    (ATen/NumericUtils.h:149-160: lg2.approx * ln2), which no CPU log reproduces bit for bit;
    bit-exact checks of this transform are done on the GPU against torch itself."""
    u = u32_to_uniform(raw_u32(numel, seed, offset, sm_count))
    eps = np.finfo(np.float32).eps
    lg = np.where(u >= np.float32(1.0) - eps / 2, np.float32(-eps / 2), np.log(u).astype(np.float32))
    return (-lg).astype(np.float32)
