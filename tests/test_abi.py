"""The C-ABI boundary without a GPU: libmvsgi_b200.so loads, exports every symbol include/mvsgi_b200.h declares, the
ctypes mirror (mvs_gi_b200/abi.py) has the same struct layouts as the C compiler, and argument errors come back as
MVSGI_E_* codes (no compute call is made here)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from mvs_gi_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "mvsgi_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int64_t|int32_t|int)\s+(mvsgi_\w+)\s*\(", text, flags=re.M)
    assert len(names) >= 10
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()  # raises if the .so is missing or its ABI version differs
    names = declared_functions()
    assert sorted(abi.EXPORTS) == names, "abi.EXPORTS and the header disagree"
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.mvsgi_abi_version() == abi.ABI_VERSION
    m = re.search(r"#define MVSGI_ABI_VERSION (\d+)", HEADER.read_text())
    assert int(m.group(1)) == abi.ABI_VERSION


def test_ctypes_structs_match_the_c_layout(tmp_path):
    structs = ["MvsgiCamera", "MvsgiRig", "MvsgiSweepArgs", "MvsgiProjectArgs", "MvsgiSoftArgminArgs",
               "MvsgiResampleArgs"]
    probes = {"MvsgiSweepArgs": ["feats", "masks", "grids", "rig", "vol", "workspace", "cand_offset", "peer_vols", "octet_order",
                                 "cta_times_out", "tap_table", "tap_table_bytes", "tap_table_mode"],
              "MvsgiSoftArgminArgs": ["costs", "inv_dist_idx", "argmax"], "MvsgiRig": ["cams", "dist"],
              "MvsgiCamera": ["rot", "p"], "MvsgiProjectArgs": ["rig", "grid_masks"],
              "MvsgiResampleArgs": ["img", "img_stride", "rig", "invalid_value", "out", "valid_out"]}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for s in structs:
        src.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in probes[s]:
            src.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    src.append("return 0; }")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(c)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        key, val = line.split()
        if "." in key:
            s, f = key.split(".")
            assert getattr(getattr(abi, s), f).offset == int(val), key
        else:
            assert C.sizeof(getattr(abi, key)) == int(val), key


def test_argument_errors_are_codes_not_crashes():
    lib = abi.load_library()
    assert lib.mvsgi_sweep_f32(None, None) == -1  # MVSGI_E_NULL
    a = abi.MvsgiSweepArgs()
    assert lib.mvsgi_sweep_workspace_bytes(C.byref(a)) == -6  # struct_size 0: MVSGI_E_VERSION
    a.struct_size = C.sizeof(a)
    assert lib.mvsgi_sweep_workspace_bytes(C.byref(a)) == -2  # batch 0: MVSGI_E_SHAPE
    s = abi.MvsgiSoftArgminArgs()
    s.struct_size = C.sizeof(s)
    assert lib.mvsgi_softargmin_f32(C.byref(s), None) == -1
    assert b"NULL" in lib.mvsgi_strerror(-1)
    assert lib.mvsgi_mask_pitch_words(2048) == 66
    with pytest.raises(RuntimeError, match="shape"):
        abi.check(-2, "x")


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under mvs_gi_b200/ may import or load it."""
    for py in (ROOT / "mvs_gi_b200").rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), py
        assert "libmvsgi_oracle" not in text, py
    for src in (ROOT / "mvs_gi_b200" / "csrc").glob("*"):
        assert not re.search(r"#include.*oracle", src.read_text()), src
