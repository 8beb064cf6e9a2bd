"""Generates tests/golden/detect_reference.npz by executing the REFERENCE'S OWN `detect(sess, model, names, image, path)`
(/root/reference/detect.py:56-88), compiled from the file as it lies, with its own `utils.postprocess.non_max_suppress`
(loaded by path) and stand-ins for the things around it that need TensorFlow / a display: `sess.run` returns the prepared
head outputs, `tf.check_numerics` passes through, `read_image` returns a blank PIL image of the wanted size, and the
matplotlib axes record what is drawn: every `patches.Rectangle(xy, w, h, linewidth=, edgecolor=)` and `ax.annotate(text, xy)`.
What this pins: the selection after NMS (argmax class, first maximum; kept iff score > threshold), the cell -> pixel
scaling with the ORIGINAL image size (detect.py:72), the box the reference draws, and the label text.
Run once, here:   python tests/golden/make_detect_golden.py"""
import ast
import importlib.util
import itertools
import os
import types

import numpy as np
from PIL import Image

REF = "/root/reference/detect.py"
REF_PP = "/root/reference/utils/postprocess.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_detect(record, head, image_size, threshold, threshold_iou):
    spec = importlib.util.spec_from_file_location("ref_postprocess", REF_PP)
    pp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pp)
    tree = ast.parse(open(REF).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "detect"][0]

    class Axes(object):
        def imshow(self, img): record["shown"] = np.asarray(img).shape
        def add_patch(self, p): record["rects"].append(p)
        def annotate(self, text, xy, color=None): record["texts"].append((text, np.array(xy, dtype=np.float64), color))
        def set_xticks(self, t): pass
        def set_yticks(self, t): pass
    fig = types.SimpleNamespace(gca=lambda: Axes(), canvas=types.SimpleNamespace(set_window_title=lambda t: record.__setitem__("title", t)))
    plt = types.SimpleNamespace(figure=lambda: fig, rcParams={"axes.prop_cycle": [{"color": "C%d" % i} for i in range(10)]})
    patches = types.SimpleNamespace(Rectangle=lambda xy, w, h, linewidth=None, edgecolor=None, facecolor=None:
                                    (np.array(xy, dtype=np.float64), float(w), float(h), float(linewidth), edgecolor))
    tf = types.SimpleNamespace(check_numerics=lambda t, name: t)
    ns = {"np": np, "plt": plt, "patches": patches, "itertools": itertools, "tf": tf,
          "utils": types.SimpleNamespace(postprocess=pp),
          "args": types.SimpleNamespace(preprocess="identity", threshold=threshold, threshold_iou=threshold_iou),
          "identity": lambda x: x,
          "read_image": lambda path: Image.new("RGB", image_size)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["detect"]


def run_case(conf, xy_min, xy_max, cells_wh, image_size, net_size, threshold=0.3, threshold_iou=0.4):
    """conf [cells, A, C], xy_min / xy_max [cells, A, 2] (cell units).  Returns (conf after the in-place NMS, drawn boxes)."""
    record = {"rects": [], "texts": []}
    conf = conf.copy()
    head = [conf[None], xy_min[None], xy_max[None]]
    detect = load_detect(record, head, image_size, threshold, threshold_iou)
    op = types.SimpleNamespace(name="t")
    model = types.SimpleNamespace(conf=types.SimpleNamespace(op=op), xy_min=types.SimpleNamespace(op=op), xy_max=types.SimpleNamespace(op=op),
                                  cell_width=cells_wh[0], cell_height=cells_wh[1])
    sess = types.SimpleNamespace(run=lambda tensors, feed_dict=None: head)
    class Placeholder(object):                                   # hashable: the reference uses it as a feed_dict key
        def get_shape(self):
            return types.SimpleNamespace(as_list=lambda: [1, net_size[1], net_size[0], 3])
    image = Placeholder()
    names = ["n%d" % i for i in range(conf.shape[-1])]
    detect(sess, model, names, image, "unused.jpg")
    assert record["title"] == "%d objects detected" % len(record["rects"])
    rect = np.array([[r[0][0], r[0][1], r[1], r[2], r[3]] for r in record["rects"]], dtype=np.float64).reshape(-1, 5)
    cls = np.array([int(t[0].split(" ")[0][1:]) for t in record["texts"]], dtype=np.int64)
    pct = np.array([t[0].split("(")[1].rstrip("%)") for t in record["texts"]])
    return head[0][0], rect, cls, pct


def main():
    rs = np.random.RandomState(41)
    out = {"numpy_version": np.array(np.__version__)}
    anchors = np.array([[1.08, 1.19], [3.42, 4.41], [6.63, 11.38], [9.42, 5.11], [16.62, 10.52]])
    for name, (cw, ch, C, K, image_size) in {"voc13": (13, 13, 20, 60, (640, 480)), "coco7": (7, 7, 80, 70, (1000, 600)),
                                              "rect": (7, 5, 3, 25, (333, 517)), "none": (5, 5, 4, 0, (100, 100))}.items():
        cells, A = cw * ch, 5
        cx = (np.arange(cells) % cw)[:, None] + rs.uniform(0, 1, size=(cells, A))
        cy = (np.arange(cells) // cw)[:, None] + rs.uniform(0, 1, size=(cells, A))
        wh = anchors[None] * np.exp(rs.normal(0, 0.5, size=(cells, A, 2)))
        xy_min = np.stack([cx, cy], -1) - wh / 2
        xy_max = np.stack([cx, cy], -1) + wh / 2
        conf = rs.uniform(0, 0.29, size=(cells, A, C))
        idx = rs.choice(conf.size, size=K, replace=False)
        conf.reshape(-1)[idx] = rs.uniform(0.3, 1.0, size=K)
        if K:                                                       # a box whose two best classes tie exactly: argmax takes the first
            conf.reshape(cells * A, C)[idx[0] // C, :2] = 0.75
        conf, xy_min, xy_max = conf.astype(np.float32), xy_min.astype(np.float32), xy_max.astype(np.float32)
        conf_out, rect, cls, pct = run_case(conf, xy_min, xy_max, (cw, ch), image_size, (cw * 32, ch * 32))
        out[name + "_conf_in"], out[name + "_xy_min"], out[name + "_xy_max"] = conf, xy_min, xy_max
        out[name + "_meta"] = np.array([cw, ch, image_size[0], image_size[1]])
        out[name + "_conf_out"], out[name + "_rect"], out[name + "_cls"], out[name + "_pct"] = conf_out, rect, cls, pct
        print(name, "kept", len(rect), "of", K, "candidates")
    np.savez_compressed(os.path.join(HERE, "detect_reference.npz"), **out)
    print("wrote detect_reference.npz %.0f KiB" % (os.path.getsize(os.path.join(HERE, "detect_reference.npz")) / 1024))


if __name__ == "__main__":
    main()
