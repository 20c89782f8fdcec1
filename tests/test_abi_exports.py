"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/veloslam_b200.h declares, and the ctypes mirrors match the C struct layouts.
No compute calls here (no GPU in this container)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "veloslam_b200.h")


@pytest.fixture(scope="module")
def lib():
    from veloslam_b200.build import build_library
    build_library()
    from veloslam_b200 import capi
    return capi.load_library()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"VS_API\s+[\w\s\*]+?\b(vs_\w+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    from veloslam_b200 import capi
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.vs_version().decode().endswith("sm_100a")


def test_ctypes_struct_layouts_match_the_header(lib, tmp_path):
    from veloslam_b200 import capi
    prog = tmp_path / "sizes.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "veloslam_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vs_laser_corr),'
        ' sizeof(vs_filters), sizeof(vs_carry), sizeof(vs_frame), sizeof(vs_result),'
        ' offsetof(vs_result, carry_out), offsetof(vs_frame, laser_counts),'
        ' offsetof(vs_carry, frames_closed), sizeof(vs_frame_rows), sizeof(vs_layout),'
        ' offsetof(vs_frame_rows, row_laser), offsetof(vs_layout, rows)); return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.LaserCorr), C.sizeof(capi.Filters), C.sizeof(capi.Carry),
            C.sizeof(capi.Frame), C.sizeof(capi.Result), capi.Result.carry_out.offset,
            capi.Frame.laser_counts.offset, capi.Carry.frames_closed.offset,
            C.sizeof(capi.FrameRows), C.sizeof(capi.Layout), capi.FrameRows.row_laser.offset,
            capi.Layout.rows.offset]
    assert got == want


def test_header_is_plain_c(tmp_path):
    prog = tmp_path / "c89.c"
    prog.write_text('#include "veloslam_b200.h"\nint main(void){vs_carry c; vs_carry_init(&c); return 0;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-c", "-I",
                           os.path.join(ROOT, "include"), str(prog), "-o", str(tmp_path / "c.o")])


def test_carry_init_matches_unload_data(lib):
    from veloslam_b200 import capi
    c = capi.carry_init()
    assert (c.last_azimuth, c.firing_skip, c.frame_meta_inited, c.is_hdl64) == (-1, 0, 0, 0)
    assert c.frame_timestamp_us == capi.VS_TIME_NONE and c.frame_skips == -1


def test_create_without_a_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.vs_create(0, 1024, 16, 1, C.byref(h))
    assert rc == 5 and not h.value        # VS_ERR_NO_DEVICE: no CPU fallback


def test_product_never_imports_the_oracle():
    """The product path must not import, link, load or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "veloslam_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|libvelo_oracle|velo_oracle\.h|oracle[/\\]_ref",
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(src), (dirpath, f)
