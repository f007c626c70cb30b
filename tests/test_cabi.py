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
