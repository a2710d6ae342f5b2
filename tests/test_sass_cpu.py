"""CPU suite: what the built library contains, read from its SASS with cuobjdump (no GPU needed).  Guards the claims DESIGN.md
makes about the kernels — sm_100a only, tcgen05 MMAs with CTA pairs (UTCHMMA.2CTA; UTCQMMA / UTCIMMA for the e4m3 / int8
paths), TMA tensor loads (UTMALDG), TMEM loads / stores (LDTM / STTM), bulk copies in the adaLN kernel (UBLKCP), the
programmatic-dependent-launch pair (PREEXIT / ACQBULK) — against a silent rebuild for another architecture or a code path
that degrades to mma.sync."""
import re
import shutil
import subprocess

import pytest

import qie_b200

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump (CUDA toolkit)")


@pytest.fixture(scope="module")
def sass():
    """{demangled-ish kernel name: SASS text} of every kernel in libqie.so"""
    if not qie_b200.LIB_PATH.exists():
        pytest.skip("libqie.so not built")
    txt = subprocess.run(["cuobjdump", "-sass", str(qie_b200.LIB_PATH)], check=True, capture_output=True, text=True).stdout
    out, name, buf = {}, None, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                out[name] = "\n".join(buf)
            name, buf = m.group(1), []
        elif name:
            buf.append(ln)
    if name:
        out[name] = "\n".join(buf)
    return out


def kernels(sass, needle):
    return {k: v for k, v in sass.items() if needle in k}


def test_only_sm_100a_code_objects():
    if not qie_b200.LIB_PATH.exists():
        pytest.skip("libqie.so not built")
    elf = subprocess.run(["cuobjdump", "-lelf", str(qie_b200.LIB_PATH)], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elf))
    assert archs == {"sm_100a"}, archs


def test_gemm_kernels_issue_tcgen05_from_tma_fed_shared_memory(sass):
    gemm = kernels(sass, "gemm_kernel")
    assert len(gemm) >= 20
    for name, text in gemm.items():
        assert re.search(r"UTC[HQI]MMA", text), name                  # tcgen05.mma kind::f16 / f8f6f4 / i8
        assert "UTMALDG.2D" in text and "LDTM" in text, name          # TMA operand loads, TMEM accumulator reads
        assert "HMMA" not in text.replace("UTCHMMA", ""), name        # no mma.sync fallback
    pair = [t for n, t in gemm.items() if "UTCHMMA.2CTA" in t]
    assert pair and all("UTCBAR.2CTA.MULTICAST" in t for t in pair)  # cta_group::2 MMAs, commit multicast to both CTAs
    assert any("UTCQMMA" in t for t in gemm.values()) and any("UTCIMMA" in t for t in gemm.values())   # e4m3 and int8 paths


def test_attention_kernel_keeps_p_in_tmem_and_uses_packed_math(sass):
    att = kernels(sass, "attn_pair_kernel")
    assert att
    for name, text in att.items():
        assert "UTCHMMA.2CTA" in text and "UTMALDG.2D.2CTA" in text, name
        assert "LDTM" in text and "STTM" in text, name                # S out of TMEM, P back into TMEM
        assert "MUFU.EX2" in text and "FADD2" in text, name
        # the online-softmax form carries the running max (FMNMX3) and the scale/subtract FFMA2; the bounded-score form
        # (template flag 1: q pre-scaled, p = 2^s) has no max instruction at all
        bounded = "ELb1EE" in name
        assert ("FMNMX3" in text) != bounded, name
        assert "HMMA" not in text.replace("UTCHMMA", ""), name


def test_adaln_kernels_use_bulk_copies_and_pdl_pair_is_present(sass):
    for needle in ("ln_mod_cta_kernel", "ln_mod_stream_kernel"):      # CTA-row form (default) and the warp-per-row ring
        ln = kernels(sass, needle)
        assert ln and all("UBLKCP" in t for t in ln.values()), needle
    for needle in ("ln_mod_cta_kernel", "ln_mod_stream_kernel", "gemm_kernel", "attn_pair_kernel"):
        for name, text in kernels(sass, needle).items():
            assert "PREEXIT" in text and "ACQBULK" in text, name      # griddepcontrol.launch_dependents / .wait


def test_eight_epilogue_warp_gemms_exist_and_do_not_spill_in_their_chunk_loop(sass):
    """The short-K 8-bit GEMMs run eight epilogue warps (template parameter EW = 8: ...ILi256E<EPI><QT>ELi2ELi8E).  Ten warps cap
    them at 168 registers, and spill reloads miss the few KB of L1 beside the operand ring (the ncu finding of
    profiles/r02_q8_qkv_ncu_stalls.txt: eight dependent LDL per 32-column chunk), so the spill traffic of these kernels is held
    to the handful of per-tile values ptxas parks today; the bf16 kernels keep four epilogue warps."""
    gemm = kernels(sass, "gemm_kernel")
    ew8 = {n: t for n, t in gemm.items() if n.endswith("ELi8EEEv14CUtensorMap_stS1_S1_NS_7GemmDevE")}
    assert len(ew8) == 20                                                # BN 256 x 5 epilogues x {e4m3, int8} x {single CTA, pair}
    for name, text in ew8.items():
        assert re.search(r"UTC[QI]MMA", text) and "UTCHMMA" not in text, name       # 8-bit operand types only
        assert len(re.findall(r"\bLDL", text)) <= 20, (name, len(re.findall(r"\bLDL", text)))
    for name, text in gemm.items():
        if "UTCHMMA" in text:                                           # bf16: four epilogue warps, at most the 3 reloads of today
            assert "ELi4EEEv14CUtensorMap" in name and len(re.findall(r"\bLDL", text)) <= 4, name
