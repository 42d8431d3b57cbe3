"""CPU tests of the host side: the C-ABI library loads without a GPU and exports every symbol the
header declares, struct layouts agree, generated header is current, synthetic inputs are stable."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from zbot_lab_b200 import native
    lib = native.lib()
    hdr = open(os.path.join(ROOT, "include", "zbot_b200.h")).read()
    declared = set(re.findall(r"\b(zbot_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(native.EXPORTED_SYMBOLS), declared ^ set(native.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.zbot_abi_version() == native.ZBOT_ABI_VERSION
    assert b"sm_100a" in lib.zbot_build_info()


def test_cfg_struct_matches_c_defaults():
    from zbot_lab_b200 import native
    lib = native.lib()
    c = native.ZbotCfg()
    assert lib.zbot_default_cfg(C.byref(c), 4096) == 0
    assert bytes(c) == bytes(native.make_cfg(4096))     # same layout, same values, same term order
    assert c.num_terms == 13 and abs(c.term_weight[0] - 0.02) < 1e-9


def test_state_field_tables_match_library():
    from zbot_lab_b200 import native, stepper
    lib = native.lib()
    used = np.zeros(80, int)
    for k, w in stepper.STATE_FIELDS.items():
        w0 = lib.zbot_state_word(k.encode())
        assert w0 >= 0, k
        used[w0:w0 + w] += 1
    assert used.max() == 1
    used = np.zeros(72, int)
    for k, w in stepper.MDP_STATE_FIELDS.items():
        w0 = lib.zbot_mdp_state_word(k.encode())
        assert w0 >= 0, k
        used[w0:w0 + w] += 1
    assert used.max() == 1
    assert lib.zbot_state_word(b"nope") == -1


def test_generated_header_is_current():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "gen_model_header.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert open(gen.HEADER).read() == gen.render()


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "zbot_lab_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(d, f)
                assert "cpu_port" not in src or f.endswith(".h") is False and "cpu_port" not in src, f


def test_env_origin_grid_matches_il_semantics():
    from oracle.il_semantics import env_origins_grid as ref
    from zbot_lab_b200.utils.synthetic import env_origins_grid
    for n in (1, 2, 7, 64, 4096):
        assert np.array_equal(env_origins_grid(n), ref(n))
    g = env_origins_grid(4096)
    assert g.shape == (4096, 3) and g[:, 0].max() == 126.0 and g[:, 1].min() == -126.0


def test_no_cuda_means_loud_failure():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zbot_lab_b200.stepper import NativeStepper
    with pytest.raises(RuntimeError):
        NativeStepper(4, "cpu")
