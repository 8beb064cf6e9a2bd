"""CPU (-m "not gpu"): the C-ABI library builds, loads and exports every symbol include/yolo2_b200.h
declares (no compute calls without a GPU); host-side logic of the Python mirror."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "yolo2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(y2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import ctypes
    from yolo_tf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = _header_functions()
    assert len(names) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "library does not export %s" % n
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert _lib.lib().y2_version() >= 100


def test_product_path_never_imports_oracle():
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "yolo_tf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt:
                    bad.append(f)
    assert not bad, bad


def test_no_cpu_fallback_on_cpu_tensor():
    import torch
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.model.yolo2 import inference
    with pytest.raises(_lib.Y2Error):
        inference.darknet(torch.zeros(1, 32, 32, 3), 20, 5)
    with pytest.raises(_lib.Y2Error):
        _lib.ptr(torch.zeros(4))


def test_calc_cell_xy():
    from yolo_tf_b200.model.yolo import calc_cell_xy
    g = calc_cell_xy(3, 5)
    assert g.shape == (3, 5, 2) and g.dtype == np.float32
    for y in range(3):
        for x in range(5):
            assert g[y, x].tolist() == [x, y]


def test_downsampling_and_config_glue(tmp_path):
    import configparser
    from yolo_tf_b200 import utils
    from yolo_tf_b200.model.yolo2 import inference
    assert inference.DARKNET_DOWNSAMPLING == (32, 32)
    cfg = configparser.ConfigParser()
    a, b = tmp_path / "a.ini", tmp_path / "b.ini"
    a.write_text("[config]\nmodel = yolo2\nbasedir = /tmp/x\n[yolo2]\ninference = darknet\nwidth = 416\n[cache]\nnames = config/names/20\n")
    b.write_text("[yolo2]\nwidth = 608\n")
    utils.load_config(cfg, [str(a), str(b)])           # later files override earlier ones (README.md:21)
    assert cfg.getint("yolo2", "width") == 608
    assert utils.calc_cell_width_height(cfg, 416, 608) == (13, 19)
    with pytest.raises(AssertionError):
        utils.calc_cell_width_height(cfg, 400, 416)
    assert utils.get_cachedir(cfg) == "/tmp/x/cache/20"


def test_variable_store_names_and_shapes():
    from yolo_tf_b200 import variables as V
    s = V.VariableStore()
    with pytest.raises(Exception):
        s.get("yolo2_darknet/conv0/weights", (3, 3, 3, 32), V.xavier_uniform, "cuda")   # no GPU here -> loud


def test_tiny_host_geometry_matches_oracle_table_and_config_dispatch(tmp_path):
    """config/yolo2/tiny-20.ini: `inference = tiny` -> inference.tiny / TINY_DOWNSAMPLING (model/yolo2/__init__.py:107,
    utils/__init__.py:47-49)."""
    import configparser
    import torch
    from oracle.darknet_oracle import tiny_layer_table
    from yolo_tf_b200 import _lib, utils, variables
    from yolo_tf_b200.model.yolo2 import inference
    host = inference.tiny_layer_geometry(20, 5)
    assert [(n, k, ci, co) for n, k, ci, co, _, _ in host] == [(n, k, ci, co) for n, k, ci, co, _ in tiny_layer_table(20, 5)]
    assert [p for *_, p in host] == [True] * 5 + ['s1', False, False, False]
    cfg = configparser.ConfigParser()
    ini = tmp_path / "tiny.ini"
    ini.write_text("[config]\nmodel = yolo2\n[yolo2]\ninference = tiny\n")
    utils.load_config(cfg, [str(ini)])
    assert utils.calc_cell_width_height(cfg, 416, 416) == (13, 13)
    assert getattr(inference, cfg.get("yolo2", "inference")) is inference.tiny
    with pytest.raises(_lib.Y2Error):
        inference.tiny(torch.zeros(1, 32, 32, 3), 20, 5)           # no CPU path
    w = variables.truncated_normal_01((3, 3, 16, 32))
    assert np.abs(w).max() <= 0.2 and 0.07 < w.std() < 0.1


def test_label_encoder_has_no_cpu_path():
    from yolo_tf_b200 import _lib
    from yolo_tf_b200.utils import data
    with pytest.raises(_lib.Y2Error):
        data.transform_labels_batch([np.array([1])], [np.zeros((1, 4), np.float32)], 20, 13, 13, device="cpu")
