"""CPU (-m "not gpu"): the drop-in aliases of INTEGRATION.md section 1 -- the reference's drivers find the hot path through module
paths (`importlib.import_module('model.' + model)` detect.py:96 / train.py:99; `utils.postprocess.non_max_suppress` detect.py:71;
`importlib.import_module('model.<name>.inference')` utils/__init__.py:47-49); `yolo_tf_b200.dropin.install()` must make exactly
those lookups land on this package's module objects.  Run in subprocesses so that sys.modules of the test process stays clean."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_aliases_resolve_the_reference_drivers_lookups(tmp_path):
    out = _run("""
        import importlib, sys
        from yolo_tf_b200 import dropin
        names = dropin.install()
        yolo = importlib.import_module('model.' + 'yolo2')                              # detect.py:96, train.py:99
        import yolo_tf_b200.model.yolo2 as ours
        assert yolo is ours and yolo.Builder is ours.Builder and yolo.Model is ours.Model and yolo.Objectives is ours.Objectives
        inf = importlib.import_module('.'.join(['model', 'yolo2', 'inference']))        # utils/__init__.py:48
        assert getattr(inf, 'darknet'.upper() + '_DOWNSAMPLING') == (32, 32)           # utils/__init__.py:49
        assert inf.darknet.__module__ == 'yolo_tf_b200.model.yolo2.inference'           # the SAME module, not a second copy
        from model.yolo2.function import reorg
        from model.yolo.function import leaky_relu
        import utils.postprocess, utils.preprocess, utils.data
        from yolo_tf_b200.utils import postprocess as pp
        assert utils.postprocess.non_max_suppress is pp.non_max_suppress                # detect.py:71
        assert utils.get_downsampling.__module__ == 'yolo_tf_b200.utils'                # no other `utils` here: ours stands in
        dropin.uninstall()
        assert 'model' not in sys.modules and 'utils.postprocess' not in sys.modules
        print('ALIASES', len(names))
    """, str(tmp_path))
    assert "ALIASES 10" in out


def test_an_existing_utils_package_keeps_its_own_functions_and_gets_the_three_submodules_swapped(tmp_path):
    """With the reference's repository on sys.path its own `utils` (config glue, drawing, ...) stays; only postprocess / preprocess /
    data are replaced.  A stand-in package plays that role here (the real one imports TensorFlow at the top)."""
    pkg = tmp_path / "utils"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("def get_logdir(config):\n    return 'theirs'\n")
    (pkg / "postprocess.py").write_text("def non_max_suppress(*a):\n    raise RuntimeError('the slow one')\n")
    out = _run("""
        import sys
        sys.path.insert(0, '.')
        from yolo_tf_b200 import dropin
        dropin.install()
        import utils, utils.postprocess
        from yolo_tf_b200.utils import postprocess as pp
        assert utils.get_logdir(None) == 'theirs'
        assert utils.postprocess.non_max_suppress is pp.non_max_suppress and sys.modules['utils.postprocess'] is pp
        print('PATCHED')
    """, str(tmp_path))
    assert "PATCHED" in out
