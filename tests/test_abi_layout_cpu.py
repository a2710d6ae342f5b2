"""CPU suite: the ctypes mirrors in qie_b200/_lib.py against the C compiler's own view of include/qie.h — size of every
struct and offset of every field, printed by a C99 program gcc builds from the header at test time.  A field added, removed
or re-ordered on one side only would corrupt every call through the ABI silently; this makes it a test failure."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import pytest

from qie_b200 import _lib as L

ROOT = Path(__file__).resolve().parent.parent
PAIRS = [("qie_model_cfg", L.ModelCfg), ("qie_seq", L.Seq), ("qie_sp", L.Sp), ("qie_block_weights", L.BlockWeights),
         ("qie_weights", L.Weights), ("qie_gemm_args", L.GemmArgs), ("qie_peers", L.Peers)]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_ctypes_mirrors_match_the_header_field_by_field(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "qie.h"', "int main(void) {"]
    for cname, mirror in PAIRS:
        lines.append(f'  printf("{cname} . %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    # -pedantic C99: the header must stay a plain C ABI (no C++-only constructs outside the extern "C" guard)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = {}
    for ln in out.splitlines():
        cname, fname, val = ln.split()
        seen[(cname, fname)] = int(val)
    for cname, mirror in PAIRS:
        assert seen[(cname, ".")] == C.sizeof(mirror), (cname, seen[(cname, ".")], C.sizeof(mirror))
        for fname, _ in mirror._fields_:
            assert seen[(cname, fname)] == getattr(mirror, fname).offset, (cname, fname)


def test_every_header_struct_has_a_mirror():
    import re
    text = (ROOT / "include" / "qie.h").read_text()
    declared = set(re.findall(r"typedef struct (\w+) \{", text))
    assert declared == {c for c, _ in PAIRS}, declared ^ {c for c, _ in PAIRS}
