"""CPU suite: the built library really is sm_100a tensor-core / TMA code, and the per-instruction findings of round 2 stay fixed.
Reads `cuobjdump -sass` of the in-tree .so (no GPU needed): tcgen05.mma shows up as UTCHMMA (`.2CTA` for cta_group::2), tcgen05.ld as
LDTM, TMA tile loads as UTMALDG, cp.async as LDGSTS (mnemonics: /opt/skills/guides/B200_PROFILING.md)."""
import os
import re
import shutil
import subprocess
from collections import Counter, defaultdict

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "crb-active-3ddet_b200", "lib", "libcrb3d_sm100.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is missing")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = defaultdict(Counter)
    name = ""
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            per[name][m.group(1)] += 1
    assert "arch = sm_100a" in txt or "sm_100a" in txt
    return per


def _kernels(per, key):
    """Kernels whose (mangled) FUNCTION name is `key`: <length><key> followed by I (template arguments) or E / v (plain function) -
    the anonymous namespace's mangling also carries the source file name."""
    pat = re.compile(r"%d%s[IEv]" % (len(key), re.escape(key)))
    return {k: v for k, v in per.items() if pat.search(k)}


def _count(c, prefix):
    return sum(n for op, n in c.items() if op.startswith(prefix))


def test_tensor_core_kernels_use_tcgen05_tmem_tma(sass):
    for key, two_cta in (("bev_conv3x3_pair_tc", True), ("bev_gemm_pair_tc", True), ("bev_gemm_tc", False), ("spconv_fwd_tc", False),
                         ("fc_gemm_tc", False)):
        ks = _kernels(sass, key)
        assert ks, key
        for name, c in ks.items():
            assert _count(c, "UTCHMMA") > 0, name            # tcgen05.mma
            assert _count(c, "LDTM") > 0, name               # tcgen05.ld (accumulators come back from tensor memory)
            assert _count(c, "UTMALDG") > 0, name            # TMA tile loads
            if two_cta:
                assert _count(c, "UTCHMMA.2CTA") > 0, name   # cta_group::2: one MMA over a CTA pair
    for name, c in _kernels(sass, "spconv_fwd_tc").items():
        assert _count(c, "LDGSTS") > 0, name                  # the row gather is cp.async
    assert not any(_count(c, "HMMA") and not _count(c, "UTCHMMA") for c in sass.values())   # no mma.sync / wmma kernels


def test_round2_instruction_level_findings_stay_fixed(sass):
    tc = {}
    for key in ("bev_conv3x3_pair_tc", "bev_conv3x3_tc", "bev_gemm_pair_tc", "bev_gemm_tc", "spconv_fwd_tc", "spconv_fwd_tc_grp", "fc_gemm_tc"):
        tc.update(_kernels(sass, key))
    assert len(tc) >= 20
    for name, c in tc.items():
        # staging and work lists are addressed in the shared window: no generic 128-bit loads / stores
        assert _count(c, "LD.E.128") == 0 and _count(c, "ST.E.128") == 0, name
    for key in ("bev_conv3x3_pair_tc", "bev_gemm_pair_tc"):
        for name, c in _kernels(sass, key).items():
            # cluster barriers (setup / teardown) carry the only cluster-scope fences; the epilogue's remote arrive has none
            assert _count(c, "MEMBAR.ALL.GPU") <= 5, (name, c["MEMBAR.ALL.GPU"])
