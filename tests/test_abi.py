"""The C-ABI library loads and exports every symbol include/lmcu.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "lmcu.h")).read()
    return sorted(set(re.findall(r"\b(lmcu_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from lmcu import api
    assert os.path.exists(api.LIB_PATH), "liblmcu.so missing: run __graft_entry__.build()"
    lib = C.CDLL(api.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(api.EXPORTS) == names


def test_struct_sizes_match_header(tmp_path):
    """Every POD struct of include/lmcu.h has the same size in the ctypes mirror as gcc gives it (layout drift between
    the header and lmcu/abi.py would corrupt scene descriptions silently)."""
    import subprocess
    from lmcu import abi
    names = ["LmcuRay", "LmcuHit", "LmcuMesh", "LmcuPrimitive", "LmcuTexture", "LmcuMaterial", "LmcuLight", "LmcuCamera",
             "LmcuSceneDesc", "LmcuRenderParams", "LmcuRenderStats", "LmcuBuildOptions", "LmcuBuildStats"]
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "lmcu.h"\nint main(void) {\n' +
                   "".join('  printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for n in names:
        assert C.sizeof(getattr(abi, n)) == int(sizes[n]), n
    assert C.sizeof(abi.LmcuRay) == 64 and C.sizeof(abi.LmcuHit) == 32
    from lmcu.scenes import HIT_DTYPE, RAY_DTYPE
    assert RAY_DTYPE.itemsize == 64 and HIT_DTYPE.itemsize == 32


def test_field_offsets_match_header(tmp_path):
    """... and every field sits at the offset gcc gives it (same names in both; a reordered or retyped field that keeps the
    struct size would otherwise go unnoticed)."""
    import subprocess
    from lmcu import abi
    names = ["LmcuRay", "LmcuHit", "LmcuMesh", "LmcuPrimitive", "LmcuTexture", "LmcuMaterial", "LmcuLight", "LmcuCamera",
             "LmcuSceneDesc", "LmcuRenderParams", "LmcuRenderStats", "LmcuBuildOptions", "LmcuBuildStats"]
    pairs = [(n, f[0]) for n in names for f in getattr(abi, n)._fields_]
    src = tmp_path / "offsets.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lmcu.h"\nint main(void) {\n' +
                   "".join('  printf("%s.%s %%zu\\n", offsetof(%s, %s));\n' % (n, f, n, f) for n, f in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "offsets"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    assert len(pairs) > 100
    for n, f in pairs:
        assert getattr(getattr(abi, n), f).offset == int(got["%s.%s" % (n, f)]), (n, f)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly (LMCU_ERR_NO_DEVICE), never compute on the CPU."""
    import lmcu
    from lmcu import scenes
    if lmcu.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(lmcu.LmcuError) as e:
        lmcu.Scene(scenes.cornell_box())
    assert e.value.code == 4


def test_product_never_references_the_checkers():
    pkg = os.path.join(ROOT, "lightmetrica-v3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "lm_oracle" not in text, f
                assert "liblmref" not in text and "libhostsim" not in text, f
