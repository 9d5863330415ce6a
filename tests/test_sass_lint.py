"""Static checks on the SASS of the built library (cuobjdump; no GPU needed).

  * the tensor-core / TMA kernels really are Blackwell-native: tcgen05.mma (UTCHMMA), TMA tensor loads (UTMALDG) and
    TMEM loads (LDTM) in every GEMM variant, the fused sampler and both tcgen05 attention kernels (the attention core
    was an mma.sync kernel);
  * no `ELECT ... BRA.U.ANY` uniformisation loop around a single-thread instruction: the single-thread regions are entered through
    elect.sync (ptx::elect_one), see DESIGN.md section 3 -- with `lane == 0` nvcc wraps every tcgen05.mma / commit / TMA issue in
    such a loop (~110 cycles per 50-cycle MMA in the attention kernel's timeline).
"""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "paella_b200", "libpaella_b200.so")


@pytest.fixture(scope="module")
def census():
    if shutil.which("cuobjdump") is None or not os.path.exists(SO):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            counts[fn][m.group(1).split(".")[0]] += 1
            if m.group(1).startswith("BRA.U.ANY"):
                counts[fn]["BRA.U.ANY"] += 1
    return counts


def _kernels(census, needle):
    return {k: v for k, v in census.items() if needle in k}


@pytest.mark.parametrize("needle", ["gemm_f16_cg2_kernel", "gemm_f16_kernel", "fused_sampler", "attention_tt_kernel", "attention_tc_kernel",
                                    "vq_mlp_fused_kernel"])
def test_tensor_core_kernels_are_tcgen05_and_tma(census, needle):
    ks = _kernels(census, needle)
    assert ks, f"no kernel matching {needle} in the library"
    for name, c in ks.items():
        assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0, (name, dict(c))
        assert c["HMMA"] == 0, f"{name}: mma.sync instructions in a tcgen05 kernel"


def test_attention_fallback_is_the_only_mma_sync_kernel(census):
    hmma = [k for k, c in census.items() if c["HMMA"] > 0]
    assert hmma and all("attention_kernel" in k for k in hmma), hmma


@pytest.mark.parametrize("needle", ["gemm_f16_cg2_kernel", "gemm_f16_kernel", "fused_sampler", "attention_tt_kernel", "attention_tc_kernel",
                                    "vq_mlp_fused_kernel"])
def test_single_thread_regions_have_no_uniformisation_loops(census, needle):
    for name, c in _kernels(census, needle).items():
        # a handful remain around mbarrier arrivals / trace hooks guarded by `lane == 0`; none may sit on an MMA or TMA issue path:
        # far fewer loops than tensor-core + TMA instructions
        assert c["BRA.U.ANY"] <= 2, (name, c["BRA.U.ANY"], c["UTCHMMA"], c["UTMALDG"])
