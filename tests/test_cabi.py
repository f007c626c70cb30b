"""The C-ABI library loads without a GPU, exports every symbol the header declares, and its value
side fails loudly (no CPU fallback) when no device is present."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(h2e):
    hdr = open(os.path.join(ROOT, "include", "h2ecc_b200.h")).read()
    names = sorted(set(re.findall(r"\b(h2e_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 14
    L = h2e.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/h2ecc_b200.h but not exported"
    assert L.h2e_version() >= 1


def test_shape_side_works_on_cpu(h2e):
    sb = h2e.ScriptBuilder()
    a, b = sb.assign_w(0), sb.assign_w(1)
    sb.int_mul(a, b)
    s = h2e.Shape.from_script(h2e.FIELD_BN256_FQ, sb.words)
    assert (s.base_offset, s.range_offset, s.n_slots, s.n_perms, s.n_instr) == (19, 44, 171, 53, 3)
    assert s.n_input_cells == 4
    with pytest.raises(h2e.H2EError):
        h2e.Shape.from_script(7, sb.words)
    with pytest.raises(h2e.H2EError):
        h2e.Shape.from_script(0, [6, 2, 5, 9])  # int_mul on integers that do not exist


def test_value_side_has_no_cpu_fallback(h2e):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sb = h2e.ScriptBuilder()
    sb.assign_w(0)
    s = h2e.Shape.from_script(0, sb.words)
    with pytest.raises(h2e.H2EError) as e:
        s.run_host(h2e.pack_inputs([[1]]))
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(h2e.H2EError):
        s.run(torch.zeros((1, 2, 32), dtype=torch.uint8))


def test_header_is_plain_c_and_links_from_c(h2e, tmp_path):
    """The boundary is a C ABI: include/h2ecc_b200.h compiles as strict C99 and a C program links against the library and
    builds a shape through it (shape side only: no GPU here)."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    h2e.lib()  # make sure the library is built
    src = tmp_path / "cabi.c"
    src.write_text(
        '#include <stdio.h>\n#include "h2ecc_b200.h"\n'
        "int main(void) {\n"
        "    /* assign_w(0), assign_w(1), int_mul(0, 1) as an op-script: (opcode, nargs, args...) */\n"
        "    const uint32_t script[] = {1, 1, 0, 1, 1, 1, 6, 2, 0, 1};\n"
        "    uint64_t q[16];\n"
        "    h2e_shape* s = h2e_shape_from_script(H2E_FIELD_BN256_FQ, script, sizeof script / sizeof script[0], 0, 0);\n"
        '    if (!s) { printf("error: %s\\n", h2e_last_error()); return 2; }\n'
        "    if (h2e_shape_query(s, q)) return 3;\n"
        '    printf("%llu %llu %llu %llu\\n", (unsigned long long)q[3], (unsigned long long)q[4], (unsigned long long)q[6], (unsigned long long)h2e_records_bytes(s, H2E_REC_PRIMARY, 32));\n'
        "    h2e_shape_free(s);\n"
        "    return h2e_version() >= 1 ? 0 : 1;\n"
        "}\n")
    libdir = os.path.join(ROOT, "halo2ecc-s_b200")
    exe = tmp_path / "cabi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lh2ecc_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert [int(x) for x in out[:3]] == [19, 44, 171]  # base offset, range offset, advice cells of assign_w, assign_w, int_mul
    assert int(out[3]) == 32 * 732  # PRIMARY record bytes of one tile
