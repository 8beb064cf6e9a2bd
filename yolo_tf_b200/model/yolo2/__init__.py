"""model/yolo2/__init__.py of the reference, B200-native: ``Model`` (head decode), ``Objectives``
(4-part loss, forward + backward) and ``Builder`` -- the call surface train.py:109-112 and
detect.py:101-102 drive.  Tensors are float32 CUDA torch tensors used as device buffers; every
value is produced by a kernel in csrc/ through the C ABI (include/yolo2_b200.h)."""
import configparser
import ctypes
import os

import numpy as np

from ... import _lib
from . import inference
from .. import yolo


class Model(object):
    """Head decode -- model/yolo2/__init__.py:28-59.

    net [B, Hc, Wc, A*(5+C)].  Attributes as in the reference (shapes [B, cells, A, ...]):
    iou, offset_xy, wh, prob, areas, offset_xy_min, offset_xy_max, wh01, wh01_sqrt, coords and, when
    not training, xy, xy_min, xy_max, conf; plus cell_height, cell_width, inputs, classes, anchors.
    The detection outputs are computed eagerly by one kernel launch; the remaining attributes by a
    second launch of the same kernel on first access.
    """
    _DETECTION = ("conf", "xy_min", "xy_max")
    _LAZY = ("iou", "prob", "wh", "areas", "xy", "offset_xy", "offset_xy_min", "offset_xy_max", "coords", "wh01")
    _SHAPE = {"conf": "C", "prob": "C", "iou": None, "areas": None, "coords": 4}
    _ANCHOR_CACHE = {}

    def __init__(self, net, classes, anchors, training=False):
        import torch
        b, self.cell_height, self.cell_width, d = net.shape
        self.anchors = np.asarray(anchors)
        a = len(self.anchors)
        if d != a * (5 + classes):
            raise ValueError("net has %d channels, expected %d" % (d, a * (5 + classes)))
        self.inputs = net
        self.classes = classes
        self.training = training
        self._b, self._a = b, a
        self._cache = {}
        key = (str(net.device), self.anchors.astype(np.float32).tobytes())
        if key not in Model._ANCHOR_CACHE:          # one upload per (device, anchor set): no per-step host sync
            Model._ANCHOR_CACHE[key] = torch.as_tensor(self.anchors.astype(np.float32)).to(net.device).contiguous()
        self._anchors_dev = Model._ANCHOR_CACHE[key]
        if not training:
            self._launch(self._DETECTION)

    def _alloc(self, name):
        import torch
        cells = self.cell_height * self.cell_width
        tail = self._SHAPE.get(name, 2)
        shape = (self._b, cells, self._a) + (() if tail is None else ((self.classes,) if tail == "C" else (tail,)))
        return torch.empty(shape, dtype=torch.float32, device=self.inputs.device)

    def _launch(self, names):
        import torch
        outs = _lib.HeadOutputs()
        for n in names:
            self._cache[n] = self._alloc(n)
            setattr(outs, n, self._cache[n].data_ptr())
        _lib.check(_lib.lib().y2_head_decode(_lib.ptr(self.inputs.contiguous(), torch.float32), self._b, self.cell_height,
                                             self.cell_width, self._a, self.classes, _lib.ptr(self._anchors_dev),
                                             ctypes.byref(outs), _lib.current_stream()))

    def __getattr__(self, name):
        if name == "wh01_sqrt":
            return self.coords[..., 2:4]
        if name in Model._DETECTION or name in Model._LAZY:
            cache = self.__dict__.get("_cache", {})
            if name not in cache:
                if name in Model._DETECTION and self.__dict__.get("training", False):
                    raise AttributeError("%s is only built when training=False (model/yolo2/__init__.py:50)" % name)
                self._launch(Model._LAZY if name in Model._LAZY else Model._DETECTION)
            return self._cache[name]
        raise AttributeError(name)


class Objectives(dict):
    """Loss -- model/yolo2/__init__.py:62-94.  Keys: iou_best, iou_normal, coords, prob (0-d CUDA
    tensors, unweighted, as the reference stores them).  ``grad_inputs`` = d(sum_k hparam_k *
    objective_k)/d(model.inputs), produced by the same fused kernel (the reference leaves this to
    TF autodiff)."""
    KEYS = ("prob", "iou_best", "iou_normal", "coords")          # C-ABI order

    def __init__(self, model, mask, prob, coords, offset_xy_min, offset_xy_max, areas, hparam=None, need_grad=True):
        import torch
        dict.__init__(self)
        self.model = model
        dev = model.inputs.device
        lab = [torch.as_tensor(t).to(dev, torch.float32).contiguous() for t in (mask, prob, coords, offset_xy_min, offset_xy_max, areas)]
        self.mask, self.prob, self.coords, self.offset_xy_min, self.offset_xy_max, self.areas = lab
        b, hc, wc = model._b, model.cell_height, model.cell_width
        cells = hc * wc
        for t, per in zip(lab, (1, model.classes, 4, 2, 2, 1)):
            if t.numel() != b * cells * per:
                raise ValueError("label tensor has %d elements, expected %d" % (t.numel(), b * cells * per))
        hp = hparam or {"prob": 1.0, "iou_best": 1.0, "iou_normal": 1.0, "coords": 1.0}
        self.hparam = dict(hp)
        L = _lib.lib()
        nbytes = L.y2_loss_workspace_bytes(b, hc, wc)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        objs = torch.empty(4, dtype=torch.float32, device=dev)
        self.grad_inputs = torch.empty_like(model.inputs) if need_grad else None
        hp_c = (ctypes.c_float * 4)(*[float(hp[k]) for k in self.KEYS])
        _lib.check(L.y2_loss_fwd_bwd(_lib.ptr(model.inputs.contiguous(), torch.float32), b, hc, wc, model._a, model.classes,
                                     _lib.ptr(model._anchors_dev), *[_lib.ptr(t) for t in lab], hp_c, _lib.ptr(objs),
                                     _lib.ptr(self.grad_inputs), _lib.ptr(ws), nbytes, _lib.current_stream()))
        self._objs = objs
        for i, k in enumerate(self.KEYS):
            self[k] = objs[i]

    def total_loss(self):
        """sum of weighted objectives = tf.losses.get_total_loss() for this model (train.py:113)."""
        return sum(self[k] * self.hparam[k] for k in self.KEYS)


class Builder(object):
    """model/yolo2/__init__.py:97-119 (derives from yolo.Builder in the reference; only the v2
    behaviour is carried).  ``Builder(args, config)`` reads names / size / anchors exactly like the
    reference; ``Builder.from_values`` builds one without files."""

    def __init__(self, args, config):
        import pandas as pd
        from ...utils import get_cachedir
        section = __name__.split('.')[-1]
        self.args = args
        self.config = config
        with open(os.path.join(get_cachedir(config), 'names'), 'r') as f:
            self.names = [line.strip() for line in f]
        self.width = config.getint(section, 'width')
        self.height = config.getint(section, 'height')
        self.anchors = pd.read_csv(os.path.expanduser(os.path.expandvars(config.get(section, 'anchors'))), sep='\t').values
        self.func = getattr(inference, config.get(section, 'inference'))

    @classmethod
    def from_values(cls, names, width, height, anchors, hparam=None, inference_name='darknet'):
        self = cls.__new__(cls)
        self.args = None
        self.config = configparser.ConfigParser()
        self.config.read_dict({'yolo2_hparam': {k: str(v) for k, v in (hparam or {"prob": 1, "iou_best": 5, "iou_normal": 1, "coords": 1}).items()}})
        self.names = list(names)
        self.width, self.height = width, height
        self.anchors = np.asarray(anchors, dtype=np.float64)
        self.func = getattr(inference, inference_name)
        return self

    def __call__(self, data, training=False):
        self.scope, self.output = self.func(data, len(self.names), len(self.anchors), training=training)
        self.model = Model(self.output, len(self.names), self.anchors, training=training)

    def create_objectives(self, labels):
        section = __name__.split('.')[-1]
        hparam = {k: self.config.getfloat(section + '_hparam', k) for k in Objectives.KEYS}
        self.objectives = Objectives(self.model, *labels, hparam=hparam)
        # the reference registers hparam-weighted copies in tf.GraphKeys.LOSSES (:117-119)
        self.losses = {'weighted_' + k: self.objectives[k] * hparam[k] for k in self.objectives}
        return self.objectives

    def backward(self, allreduce=True):
        """Gradients of the total loss w.r.t. every variable -- the tf.gradients half of
        slim.learning.create_train_op (train.py:127-129).  Returns (flat float32 bucket, {variable name: view}).
        With torch.distributed initialised, the bucket is averaged over the data-parallel replicas with ONE
        all-reduce (the reference has no multi-GPU support, README.md:99)."""
        from ... import parallel
        eng = inference._Engine.get(self.output.device, len(self.names), len(self.anchors), self._arch())
        flat, views = eng.backward(self.objectives.grad_inputs)
        if allreduce:
            parallel.allreduce_mean_(flat)
        return flat, {self.scope + "/" + k: v for k, v in views.items()}

    def _arch(self):
        return inference.ARCH_TINY if self.func in (inference.tiny, inference._tiny) else inference.ARCH_DARKNET
