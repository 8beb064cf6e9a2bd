"""Variable store: the stand-in for TensorFlow's graph variables / checkpoint that the reference's
call surface assumes (``slim.variable``, ``tf.global_variables``, ``slim.assign_from_checkpoint_fn``:
detect.py:104-106, parse_darknet_yolo2.py:71).  Names and layouts are the TF ones:

    yolo2_darknet/conv{i}/weights                      [k, k, cin, cout]  (HWIO)
    yolo2_darknet/conv{i}/BatchNorm/{gamma,beta,moving_mean,moving_variance}
    yolo2_darknet/conv/{weights,biases}

Variables are float32 CUDA tensors (device-memory containers); the engine repacks them for the
tensor cores when they change.
"""
import math

import numpy as np


class VariableStore(object):
    def __init__(self):
        self._vars = {}
        self.version = 0

    def get(self, name, shape, initializer, device):
        import torch
        if name not in self._vars:
            self._vars[name] = torch.as_tensor(np.asarray(initializer(shape), dtype=np.float32)).to(device).contiguous()
            self.version += 1
        v = self._vars[name]
        if tuple(v.shape) != tuple(shape):
            raise ValueError("variable %s has shape %s, graph wants %s" % (name, tuple(v.shape), tuple(shape)))
        return v

    def assign(self, values):
        """values: dict name -> ndarray / tensor (like a checkpoint restore)."""
        import torch
        for name, val in values.items():
            # (a tensor is copied: the store must not alias a buffer the caller may overwrite without a version bump)
            t = val.detach().clone().float() if torch.is_tensor(val) else torch.as_tensor(np.asarray(val, dtype=np.float32))
            if name in self._vars:
                if tuple(self._vars[name].shape) != tuple(t.shape):
                    raise ValueError("checkpoint shape mismatch for %s" % name)
                self._vars[name].copy_(t)
            else:
                self._vars[name] = t.cuda().contiguous()
        self.version += 1

    def global_variables(self):
        return dict(self._vars)

    def __contains__(self, name):
        return name in self._vars


_default = VariableStore()


def default_store():
    return _default


def reset_default_store():
    """tf.reset_default_graph() analogue."""
    global _default
    _default = VariableStore()
    return _default


_rs = np.random.RandomState(0)


def xavier_uniform(shape):
    """slim.layers.conv2d default weights_initializer (initializers.xavier_initializer())."""
    k1, k2, cin, cout = shape
    lim = math.sqrt(6.0 / (k1 * k2 * cin + k1 * k2 * cout))
    return _rs.uniform(-lim, lim, size=shape)


def truncated_normal_01(shape):
    """tf.truncated_normal_initializer(stddev=0.1) (tiny, model/yolo2/inference.py:33): N(0, 0.1) with draws beyond two
    standard deviations re-drawn."""
    out = _rs.normal(0.0, 0.1, size=shape)
    bad = np.abs(out) > 0.2
    while bad.any():
        out[bad] = _rs.normal(0.0, 0.1, size=int(bad.sum()))
        bad = np.abs(out) > 0.2
    return out


def zeros(shape):
    return np.zeros(shape)


def ones(shape):
    return np.ones(shape)
