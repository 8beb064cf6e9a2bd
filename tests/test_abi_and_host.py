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


def test_builder_constructor_reads_names_size_anchors_and_inference_like_the_reference(tmp_path):
    """model/yolo2/__init__.py:97-107: names from <cachedir>/names, width / height / anchors (TSV with a header row,
    config/yolo2/anchors/*.tsv) and the string-dispatched inference function from the [yolo2] section."""
    import configparser
    from yolo_tf_b200.model.yolo2 import Builder, inference
    base = tmp_path / "base"
    (base / "cache" / "20").mkdir(parents=True)
    (base / "cache" / "20" / "names").write_text("\n".join("class%d" % i for i in range(20)) + "\n")
    tsv = tmp_path / "voc.tsv"
    tsv.write_text("width\theight\n1.08\t1.19\n3.42\t4.41\n6.63\t11.38\n9.42\t5.11\n16.62\t10.52\n")
    cfg = configparser.ConfigParser()
    cfg.read_dict({"config": {"basedir": str(base), "model": "yolo2"}, "cache": {"names": "config/names/20"},
                   "yolo2": {"width": "416", "height": "608", "anchors": str(tsv), "inference": "tiny"},
                   "yolo2_hparam": {"prob": "1", "iou_best": "5", "iou_normal": "1", "coords": "1"}})
    b = Builder(None, cfg)
    assert b.names == ["class%d" % i for i in range(20)] and (b.width, b.height) == (416, 608)
    assert b.anchors.shape == (5, 2) and b.anchors.dtype == np.float64 and b.anchors[2].tolist() == [6.63, 11.38]
    assert b.func is inference.tiny
    cfg.set("yolo2", "inference", "darknet")
    assert Builder(None, cfg).func is inference.darknet
    v = Builder.from_values(b.names, 416, 608, b.anchors)
    assert v.func is inference.darknet and v.config.getfloat("yolo2_hparam", "iou_best") == 5.0    # config.ini:98-102 defaults


def test_integration_md_ctypes_stub_matches_the_binding():
    """INTEGRATION.md section 2 is the binding a maintainer of the reference would paste: execute that code block as written and
    compare every argtypes / restype it sets with yolo_tf_b200/_lib.py (which the GPU suite exercises)."""
    import ctypes
    from yolo_tf_b200 import _lib
    _lib.lib()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "ctypes.CDLL" in b]
    assert len(stub) == 1
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        exec(compile(stub[0], "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    L = ns["L"]
    checked = 0
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        got = list(fn.argtypes)
        assert len(got) == len(args), (name, len(got), len(args))
        for a, b in zip(got, args):
            # c_void_p vs POINTER(struct) are both pointers; scalars must agree exactly
            pa, pb = (a is ctypes.c_void_p or hasattr(a, "contents")), (b is ctypes.c_void_p or b is ctypes.c_char_p or hasattr(b, "contents"))
            assert (pa and pb) or a is b, (name, a, b)
        if fn.restype is not ctypes.c_int:                      # ctypes' default restype is c_int
            assert fn.restype is res, (name, fn.restype, res)
        checked += 1
    assert checked >= 18


def test_binding_argument_counts_and_kinds_match_the_header():
    """A ctypes signature that disagrees with the C prototype corrupts the call silently: parse include/yolo2_b200.h and compare
    every function's parameter list (count; pointer / integer / float / size_t kind) and return type with yolo_tf_b200/_lib.py."""
    import ctypes
    from yolo_tf_b200 import _lib
    src = open(os.path.join(ROOT, "include", "yolo2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct y2_head_outputs \{.*?\} y2_head_outputs;", "", src, flags=re.S)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(y2_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) == len(_lib._SIGNATURES), (len(protos), len(_lib._SIGNATURES))

    def kind_of_c(decl):                 # (on LP64 ctypes aliases c_size_t / c_ulonglong / c_ulong: 64-bit integers form one kind)
        d = " ".join(decl.split())
        if "*" in d or "[" in d:
            return "ptr"
        if d.startswith("size_t") or d.startswith("long long") or d.startswith("unsigned long long"):
            return "i64"
        if d.startswith("float"):
            return "float"
        if d.startswith("int") or d.startswith("uint32_t"):
            return "i32"
        raise AssertionError("unparsed parameter: %r" % d)

    def kind_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents"):
            return "ptr"
        if t is ctypes.c_float:
            return "float"
        assert t in (ctypes.c_int, ctypes.c_size_t, ctypes.c_longlong, ctypes.c_ulonglong), t
        return "i32" if ctypes.sizeof(t) == 4 else "i64"

    for ret, name, params in protos:
        res, args = _lib._SIGNATURES[name]
        plist = [p.strip() for p in params.split(",")] if params.strip() not in ("", "void") else []
        assert len(plist) == len(args), (name, plist, args)
        for decl, t in zip(plist, args):
            assert kind_of_c(decl) == kind_of_ctypes(t), (name, decl, t)
        r = " ".join(ret.split())
        if r == "void":
            assert res is None, name
        elif "*" in r:
            assert res in (ctypes.c_char_p, ctypes.c_void_p), name
        else:
            assert kind_of_c(r + " x") == kind_of_ctypes(res), (name, r, res)
