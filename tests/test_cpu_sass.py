"""CPU: what the SHIPPED libkmg.so contains, read with cuobjdump (no GPU needed): code for sm_100a only; the Blackwell
instructions the design rests on are in the SASS of the Gram GEMM -- UTCIMMA.2CTA (tcgen05.mma.cta_group::2), UTMALDG.2D.2CTA
(TMA operand loads into a CTA pair), UTMASTG.2D (TMA stores of the mirrored / transposed tiles), LDTM (tcgen05.ld from
TMEM), UTCBAR.2CTA.MULTICAST (tcgen05.commit to both CTAs) -- and the headline kernel instantiations keep their registers
without a local-memory stack to speak of (B200_PROFILING.md: the mnemonics that prove tcgen05 / TMA; profiles/
r2_sass_histogram.txt holds the full histogram)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kernel-methods-for-genomics_b200", "libkmg.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(CUOBJDUMP) and os.path.exists(LIB)), reason="needs cuobjdump and the built library")


def _run(*args):
    return subprocess.run([CUOBJDUMP, *args, LIB], capture_output=True, text=True, timeout=600).stdout


def test_library_holds_sm100a_code_only():
    elfs = re.findall(r"ELF file\s+\d+:\s+(\S+)", _run("-lelf"))
    assert elfs and all(name.endswith(".sm_100a.cubin") for name in elfs), elfs
    assert "PTX file" not in _run("-lptx")  # no PTX for a JIT to retarget: sm_100a SASS or nothing


def test_gemm_sass_uses_tcgen05_tmem_and_tma():
    sass = _run("-sass", "-fun", "gram_i8_2cta_kernel") or _run("-sass")
    if "UTCIMMA" not in sass:  # older cuobjdump: -fun wants the mangled name; fall back to the whole library
        sass = _run("-sass")
    for mnemonic in ("UTCIMMA.2CTA", "UTMALDG.2D.2CTA", "UTMASTG.2D", "LDTM", "UTCBAR.2CTA.MULTICAST", "SYNCS.PHASECHK.TRANS64.TRYWAIT"):
        assert mnemonic in sass, mnemonic
    for legacy in ("HMMA", "IMMA", "HGMMA"):  # no mma.sync / wgmma path anywhere near the GEMM
        assert not re.search(r"\b%s\b" % legacy, sass), legacy


def test_headline_kernels_keep_their_state_in_registers():
    usage = _run("-res-usage")
    seen = 0
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", usage):
        name, regs, stack = m.group(1), int(m.group(2)), int(m.group(3))
        if "gram_i8_2cta_kernel" in name:
            seen += 1
            assert regs <= 168 and stack <= 16, (name, regs, stack)   # 384 threads x 168 registers = one CTA per SM, by design
        if "wd_kernel" in name and "wds" not in name:
            assert stack == 0, (name, stack)
    assert seen >= 4
