"""CPU tests of the drop-in boundary: libmms_b200.so loads, exports every symbol that
include/mms_b200.h declares, and fails loudly (no fallback) when no B200 is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from mms_answer_selection_b200 import build
    return ctypes.CDLL(build.build())


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mms_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mms_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_whole_path():
    names = declared_symbols()
    for layer in ("embed", "simcross", "simmatrix", "pairrankloss", "fm"):
        for d in ("forward", "backward"):
            for t in ("f32", "f64"):
                assert "mms_%s_%s_%s" % (layer, d, t) in names
    assert {"mms_create", "mms_destroy", "mms_set_stream", "mms_last_error", "mms_scale_f32",
            "mms_dot_f32", "mms_rerank_scores_f32"} <= set(names)


def test_library_exports_every_declared_symbol(built_lib):
    missing = [n for n in declared_symbols() if not hasattr(built_lib, n)]
    assert not missing, "declared in include/mms_b200.h but not exported: %s" % missing


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mms_b200.h"\nint main(void){ return mms_version() == MMS_B200_VERSION ? 0 : 1; }\n')
    import subprocess
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I",
                        os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    built_lib.mms_last_error.restype = ctypes.c_char_p
    assert built_lib.mms_device_ok() == 0
    h = ctypes.c_void_p()
    rc = built_lib.mms_create(ctypes.byref(h))
    assert rc != 0 and not h.value
    assert built_lib.mms_last_error()
    import mms_answer_selection_b200 as mms
    with pytest.raises(mms.MMSError):
        mms.EmbedLayer(mms.LayerParameter("Embed", embed_param=dict(num_output=4, input_dim=5)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mms_answer_selection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("the oracle", "").lower() or f == "build.py" or \
                    "import oracle" not in text and "from oracle" not in text and "mms_oracle" not in text, f
