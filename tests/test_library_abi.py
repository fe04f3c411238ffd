"""The C-ABI shared library loads in the GPU-less container and exports every symbol include/oceanostics_b200.h
declares; compute entry points refuse to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from helpers import ocb, ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "oceanostics_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build_library()
    L = ocb._lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in the header but not exported"
    assert sorted(ocb._lib.EXPORTED_SYMBOLS) == declared
    assert L.ocb_abi_version() == ocb._lib.OCB_ABI_VERSION == 2


def test_struct_layouts_match_header_sizes():
    # sizes the C compiler gives the descriptors (computed once with gcc from the header itself)
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "oceanostics_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ocb_grid), sizeof(ocb_field), sizeof(ocb_closure), sizeof(ocb_request), sizeof(ocb_sweep_inputs), sizeof(ocb_filter_pass), sizeof(ocb_ring));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    L = ocb._lib
    assert sizes == [C.sizeof(L.ocb_grid), C.sizeof(L.ocb_field), C.sizeof(L.ocb_closure), C.sizeof(L.ocb_request),
                     C.sizeof(L.ocb_sweep_inputs), C.sizeof(L.ocb_filter_pass), C.sizeof(L.ocb_ring)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    g = ocb.RectilinearGrid((4, 4, 4), extent=(1, 1, 1), device="cpu")
    m = ocb.NonhydrostaticModel(g)
    f = ocb.Field(ocb.KineticEnergyEquation.KineticEnergy(m))
    with pytest.raises(ocb.OcbError):
        f.compute()
    with pytest.raises(ocb.OcbError):
        m.velocities.u.fill_halo_regions()
    assert ocb._lib.lib().ocb_device_count() == 0
