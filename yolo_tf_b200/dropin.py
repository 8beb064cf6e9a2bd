"""Drop-in aliases: make `import model...` / `utils.postprocess` of the reference's drivers resolve to this package.

The reference's drivers reach the hot path through module paths -- `importlib.import_module('model.' + model)`
(detect.py:96, train.py:99), `utils.postprocess.non_max_suppress` (detect.py:71), `utils.get_downsampling`'s
`importlib.import_module('model.<name>.inference')` (utils/__init__.py:47-49).  `install()` registers this package's modules
in `sys.modules` under those names (the SAME module objects, not second copies: the mirror uses package-relative imports, so
aliasing only the top-level `model` name would break them), and swaps `postprocess` / `preprocess` / `data` on whatever
`utils` package is importable (the reference's own, when its repository is on sys.path) or registers this package's `utils`
when none is.
"""
import importlib
import sys

_MODEL_MODULES = ("model", "model.yolo", "model.yolo.function", "model.yolo2", "model.yolo2.inference", "model.yolo2.function")
_UTILS_MODULES = ("postprocess", "preprocess", "data")


def install(replace_utils=None):
    """Register the aliases; returns the list of module names that now resolve to this package.
    replace_utils: True = alias the whole `utils` package to ours; False = only swap the three submodules on the existing one;
    None = False if a `utils` package can be imported, else True."""
    done = []
    for name in _MODEL_MODULES:
        sys.modules[name] = importlib.import_module("yolo_tf_b200." + name)
        done.append(name)
    # attribute access `model.yolo2` on the alias keeps working because the module objects are the originals
    ours = importlib.import_module("yolo_tf_b200.utils")
    theirs = None
    if replace_utils is None or replace_utils is False:
        try:
            theirs = importlib.import_module("utils")
        except Exception:                       # not importable (absent, or its own imports fail: TensorFlow / matplotlib)
            theirs = None
        if theirs is None and replace_utils is False:
            raise ImportError("dropin.install(replace_utils=False): no importable `utils` package to patch")
    if theirs is None or theirs is ours:
        sys.modules["utils"] = ours
        done.append("utils")
        target = ours
    else:
        target = theirs
    for sub in _UTILS_MODULES:
        mod = importlib.import_module("yolo_tf_b200.utils." + sub)
        setattr(target, sub, mod)
        sys.modules["utils." + sub] = mod
        done.append("utils." + sub)
    return done


def uninstall():
    """Remove every alias that points into this package (tests)."""
    for name in list(sys.modules):
        mod = sys.modules[name]
        if (name == "model" or name.startswith("model.") or name == "utils" or name.startswith("utils.")) and \
                getattr(mod, "__name__", "").startswith("yolo_tf_b200."):
            del sys.modules[name]
