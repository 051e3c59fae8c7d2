"""No-GPU checks of the drop-in boundary: the shared library loads, exports every symbol include/p5_b200.h declares,
the ctypes mirrors match the C structs, and the product path fails loudly without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "p5_b200.h")


def _declared():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"\b(p5_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(built_lib):
    from openp5_b200 import _lib
    declared = _declared()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in p5_b200.h but not exported"
    assert sorted(_lib.DECLARED_SYMBOLS) == declared


def test_struct_layout_matches_header(built_lib, tmp_path):
    from openp5_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "p5_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(P5Config), sizeof(P5GemmDesc));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b = subprocess.check_output([str(exe)]).decode().split()
    assert int(a) == C.sizeof(_lib.P5Config)
    assert int(b) == C.sizeof(_lib.P5GemmDesc)


def test_version_and_error_string(built_lib):
    assert built_lib.p5_version() >= 100
    assert built_lib.p5_launch_count() == 0


def test_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from openp5_b200 import _lib
    cfg = _lib.P5Config()
    cfg.d_kv = 64
    h = C.c_void_p()
    rc = built_lib.p5_create(C.byref(cfg), 0, None, C.byref(h))
    assert rc != 0 and b"no CPU fallback" in built_lib.p5_last_error()
    from openp5_b200.model import P5B200
    with pytest.raises(_lib.P5LibraryError):
        P5B200("t5-small")


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under openp5_b200/ may reference it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "openp5_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8", errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "p5_oracle" not in txt, f


def test_sass_contains_blackwell_instructions(built_lib):
    from openp5_b200 import _lib
    try:
        out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "UTCHMMA" in out, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in out, "TMA loads missing from SASS"
    assert "LDTM" in out, "tcgen05.ld missing from SASS"
    assert "UTCHMMA.2CTA" in out and "UTMALDG.4D.2CTA" in out, "cta_group::2 GEMM variant missing from SASS"
    assert "HMMA.16816.F32.BF16" in out and "LDSM" in out, "mma.sync / ldmatrix decoder attention missing from SASS"
    assert "ACQBULK" in out and "PREEXIT" in out, "programmatic dependent launch (griddepcontrol) missing from SASS"


def test_gemm_tile_width_model(built_lib):
    """the tile-width cost model of the tcgen05 GEMM (gemm_tc.cu::gemm_tc_tile_width; host arithmetic, no device): the
    decisions DESIGN.md §8 / profiles/r02_session2_ab.txt describe, on 148 SMs"""
    tw = built_lib.p5_gemm_tile_width
    tw.restype = C.c_int
    tw.argtypes = [C.c_int] * 4
    # T5-base train step, 12800 packed rows: N = 768 in 2.7 waves of 192-wide tiles instead of 2.03 waves of 256-wide ones
    assert tw(12800, 768, 1, 148) == 192
    assert tw(12800, 2304, 1, 148) == 256 and tw(12800, 3072, 1, 148) == 256
    # eval encoder (20 users, ~3.9 k packed rows): N = 768 in ONE round of 192-wide tiles
    assert tw(3900, 768, 1, 148) == 192 and tw(3900, 2304, 1, 148) == 256
    # decoder rows of a train step (M = 512): every SM gets at most one tile; N = 3072 fits one round of 128-wide tiles
    assert tw(512, 768, 1, 148) == 64 and tw(512, 2304, 1, 148) == 64 and tw(512, 3072, 1, 148) == 128
    assert tw(512, 32100, 1, 148) == 256           # LM head: many column tiles, wide tile
    assert tw(400, 64, 1, 148) == 64               # narrow outputs never take a tile that is mostly padding
    assert tw(0, 768, 1, 148) == 0 and tw(128, 768, 1, 0) == 0
