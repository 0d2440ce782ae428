"""CPU tier: the C-ABI library builds, loads and exports what include/btb200.h declares;
without a GPU every compute entry point refuses loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, have_gpu

import gr_bluetooth_b200 as g


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "btb200.h")).read()
    return sorted(set(re.findall(r"BTB200_API[^;(]*?\b(btb200_\w+)\s*\(", hdr)))


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    for s in ("btb200_create", "btb200_process", "btb200_process_device", "btb200_destroy",
              "btb200_get_mm_state", "btb200_set_mm_state", "btb200_get_stage", "btb200_strerror"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    L = g.lib()
    for s in declared_symbols():
        assert hasattr(L, s), s
    assert set(g.EXPORTS) <= set(declared_symbols())


def test_struct_layouts_match_header():
    # sizes the C compiler sees for the POD structs (gcc, same ABI as nvcc's host compiler)
    import subprocess, tempfile
    src = '#include "btb200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n",sizeof(btb200_config),sizeof(btb200_info),sizeof(btb200_hit),sizeof(btb200_hits));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    assert [int(v) for v in out] == [C.sizeof(g.Config), C.sizeof(g.Info), g.HIT_DTYPE.itemsize, C.sizeof(g.Hits)]


def test_strerror_and_version():
    L = g.lib()
    assert b"no CPU fallback" in L.btb200_strerror(-2)
    assert L.btb200_strerror(0) == b"ok"
    assert "sm_100a" in g.version()


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(g.Btb200Error) as e:
        g.multi_sniffer(2e6, 2476e6, 10.0)
    assert e.value.code == -2


def test_bad_config_rejected():
    L = g.lib()
    ctx = C.c_void_p()
    cfg = g.Config(abi_version=99, sample_rate=2e6, center_freq=2476e6, squelch_threshold=10.0)
    assert L.btb200_create(C.byref(cfg), C.byref(ctx)) == -1
    assert L.btb200_create(None, C.byref(ctx)) == -1
