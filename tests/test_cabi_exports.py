"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/alad_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from aladin_b200 import build
    return build.build()


def test_header_symbols_are_exported(built_lib):
    from aladin_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "alad_b200.h")).read()
    declared = set(re.findall(r"\b(alad_[a-z0-9_]+)\s*\(", header))
    assert declared, "no entry points parsed from the header"
    assert declared == set(_cabi.PROTOTYPES), "ctypes prototypes and header disagree"
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libalad_b200.so"
    assert lib.alad_abi_version() == 5


def test_kernels_are_blackwell_native(built_lib):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # TMA
    assert "LDTM" in sass         # tcgen05.ld
    assert "sm_100a" in sass


def test_scoring_kernel_resources(built_lib):
    """The tcgen05 kernel must stay spill-free and inside the register budget of 6 warps x 1 CTA/SM; its performance is
    sensitive to code-generation changes (profiles/r01_tile_order_sweep.md sections 9, 10), so a change in these numbers is
    a reason to re-run the same-box A/B (tools/ab_prev.sh) before trusting a bench line."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-res-usage", built_lib], capture_output=True, text=True).stdout
    seen = 0
    lines = out.splitlines()
    for i, line in enumerate(lines):
        if ("mrsw_fwd_kernel" in line or "mrsw_fwd_tf32_kernel" in line) and i + 1 < len(lines):
            m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", lines[i + 1])
            assert m, lines[i + 1]
            reg, stack, _, local = map(int, m.groups())
            assert stack == 0 and local == 0, f"spills in {line.strip()}"
            assert reg <= 168, f"{reg} registers in {line.strip()}"     # 135-136 today; a jump = different code generation
            seen += 1
    assert seen == 5                                     # single-CTA, CTA-pair (bf16 and TF32 operands each) and the pair-list (two-stage) variant
