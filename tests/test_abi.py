"""CPU: the C-ABI shared library loads without a GPU, exports every symbol include/kmanip_b200.h declares, its struct
layouts agree with the ctypes mirrors, and it fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kmanip_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(km_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from gym_kmanip_b200 import _lib
    _lib.build()
    L = _lib.load()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/kmanip_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "gym_kmanip_b200/_lib.py EXPORTS must list exactly the header's functions"
    assert b"sm_100a" in L.km_version()


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of km_model, km_task, km_step_out as compiled by gcc from the header == the ctypes mirrors."""
    from gym_kmanip_b200 import _lib, flatmodel
    csrc = tmp_path / "sz.c"
    csrc.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "kmanip_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(km_model), sizeof(km_task), sizeof(km_step_out),'
                    ' offsetof(km_model, body_parent), offsetof(km_model, mocap_quat0), offsetof(km_task, q_home),'
                    ' offsetof(km_task, cube_spawn_hi)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(csrc), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(flatmodel.CModel), C.sizeof(flatmodel.CTask), C.sizeof(_lib.StepOut), flatmodel.CModel.body_parent.offset,
            flatmodel.CModel.mocap_quat0.offset, flatmodel.CTask.q_home.offset, flatmodel.CTask.cube_spawn_hi.offset]
    assert got == want


def test_oracle_structs_share_the_layout():
    """oracle/ko_model.h mirrors the same plain-C layout (the oracle is fed the identical flat model)."""
    a = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    b = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "oracle", "ko_model.h")).read(), flags=re.S)
    fa = re.search(r"typedef struct km_model \{(.*?)\} km_model;", a, re.S).group(1)
    fb = re.search(r"typedef struct ko_model \{(.*?)\} ko_model;", b, re.S).group(1)
    norm = lambda s: re.sub(r"\s+", " ", s).strip()   # noqa: E731
    assert norm(fa) == norm(fb)
    ta = re.search(r"typedef struct km_task \{(.*?)\} km_task;", a, re.S).group(1)
    tb = re.search(r"typedef struct ko_task \{(.*?)\} ko_task;", b, re.S).group(1)
    ids = lambda s: re.findall(r"\b([a-z_0-9]+)\s*(?:\[|;|,)", norm(s).replace("KM_", "KO_"))   # noqa: E731
    assert ids(ta) == ids(tb)


def test_no_cpu_fallback():
    """Without a CUDA device km_create must refuse (KM_ERR_NODEVICE) and the Python host must raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is attached")
    from gym_kmanip_b200 import _lib, constants as K, flatmodel, mjcf
    L = _lib.load()
    flat = mjcf.load_flat("solo_arm")
    pm = flatmodel.PackedModel(flat)
    task = flatmodel.make_task(flat, K.ENV_REGISTRY["KManipSoloArm"])
    h = C.c_void_p()
    rc = L.km_create(pm.ref(), C.byref(task), 0, 4, 0, 32, 0, 0, C.byref(h))
    assert rc == -4 and b"no usable CUDA device" in L.km_last_error()
    import gym_kmanip_b200 as k
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        k.make("KManipSoloArm")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        k.make_vec("KManipSoloArmQPos", 8)


def test_product_package_never_imports_the_oracle():
    """oracle/ and tests/hostsim are checkers only: nothing under gym_kmanip_b200/ may reference them."""
    pkg = os.path.join(ROOT, "gym_kmanip_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "kmanip_oracle" not in txt and "hostsim" not in txt.replace("tests/hostsim", ""), f
    out = subprocess.check_output([sys.executable, "-c", "import sys; import gym_kmanip_b200; "
                                   "print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"], cwd=ROOT)
    assert out.strip() == b"False"


def test_error_codes_and_messages_without_a_device():
    """Argument errors are reported through return codes + km_last_error (no exceptions, no aborts), also with no GPU."""
    from gym_kmanip_b200 import _lib
    L = _lib.load()
    null = C.c_void_p(None)
    assert L.km_reset(null, None, None, None, None) == -1 and b"null handle" in L.km_last_error()
    assert L.km_step(null, None, None, 0, None) == -1
    assert L.km_configure(null, 32, 0) == -1
    assert L.km_get_state(null, None, None, None, None) == -1
    assert L.km_site_poses(null, None, None, None) == -1
    h = C.c_void_p()
    assert L.km_create(None, None, 0, 4, 0, 32, 0, 0, C.byref(h)) == -1 and b"bad argument" in L.km_last_error()
    L.km_destroy(null)      # a no-op, must not crash
