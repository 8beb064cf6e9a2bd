"""Optimizer step of the training driver -- SURVEY section 8(f) row 1.

Mirrors the pieces of the reference's train.py that sit after the gradients:

    get_optimizer(config, name)                      train.py:70-80   (default name 'adam', train.py:157)
    tf.train.exponential_decay(...)                  train.py:118-124 (config.ini [exponential_decay])
    slim.learning.create_train_op(total_loss, optimizer, global_step, clip_gradient_norm=args.gradient_clip)   train.py:127-129

The update runs on the device over the flat gradient bucket that `Builder.backward` fills (after the data-parallel
all-reduce), through `y2_adam_step` (csrc/y2_optim.cu).  Only Adam -- the reference's default -- is implemented; the other
names of `get_optimizer` raise NotImplementedError.  No CPU path exists.
"""
import ctypes
import math

from . import _lib
from . import variables as V


def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
    """tf.train.exponential_decay: learning_rate * decay_rate ** (global_step / decay_steps) (floor if staircase)."""
    p = float(global_step) / float(decay_steps)
    if staircase:
        p = math.floor(p)
    return float(learning_rate) * float(decay_rate) ** p


class AdamOptimizer(object):
    """tf.train.AdamOptimizer(learning_rate, beta1, beta2, epsilon); learning_rate may be a callable of the global step
    (how train.py:120 feeds the decayed rate)."""

    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, float(beta1), float(beta2), float(epsilon)

    def rate(self, global_step):
        return float(self.learning_rate(global_step)) if callable(self.learning_rate) else float(self.learning_rate)


def get_optimizer(config, name):
    """train.py:70-80: returns a constructor taking the learning rate."""
    section = 'optimizer_' + name
    if name == 'adam':
        return lambda learning_rate: AdamOptimizer(learning_rate, config.getfloat(section, 'beta1'), config.getfloat(section, 'beta2'),
                                                   config.getfloat(section, 'epsilon'))
    raise NotImplementedError("optimizer '%s': only 'adam' (the reference's default, train.py:157) runs on the device" % name)


class TrainOp(object):
    """The callable slim.learning.create_train_op returns: one call = forward (batch statistics) + objectives + backward
    + all-reduce + per-tensor gradient clipping + Adam update + global_step increment.  Returns the total loss (device scalar)."""

    def __init__(self, builder, optimizer, global_step=0, clip_gradient_norm=0.0):
        self.builder, self.optimizer = builder, optimizer
        self.global_step = int(global_step)
        self.clip_gradient_norm = float(clip_gradient_norm)
        self._state = None

    def _prepare(self, flat, views):
        import torch
        from .model.yolo2 import inference
        L = _lib.lib()
        eng = inference._Engine.get(flat.device, len(self.builder.names), len(self.builder.anchors), self.builder._arch())
        store = V.default_store()
        names = list(views.keys())                             # bucket order: per layer weights, gamma, beta | biases
        if L.y2_num_param_tensors(eng.h) != len(names):
            raise _lib.Y2Error("optimizer: %d variables in the bucket, the engine reports %d" % (len(names), L.y2_num_param_tensors(eng.h)))
        tensors = [store.global_variables()[n] for n in names]
        for n, t, g in zip(names, tensors, views.values()):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == g.numel()):
                raise _lib.Y2Error("optimizer: variable %s is not a contiguous float32 CUDA tensor of the gradient's size" % n)
        ptrs = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        need = L.y2_adam_workspace_bytes(eng.h)
        ws = torch.empty(need + 256, dtype=torch.uint8, device=flat.device)
        self._state = {"eng": eng, "store": store, "tensors": tensors, "ptrs": ptrs, "ws": ws, "need": need,
                       "m": torch.zeros_like(flat), "v": torch.zeros_like(flat)}

    def apply_gradients(self, flat, views):
        """optimizer.apply_gradients on the (already all-reduced) bucket; increments global_step."""
        if self._state is None or self._state["m"].numel() != flat.numel() or self._state["store"] is not V.default_store():
            self._prepare(flat, views)
        st = self._state
        L = _lib.lib()
        off = (-st["ws"].data_ptr()) % 256
        lr = self.optimizer.rate(self.global_step)
        _lib.check(L.y2_adam_step(st["eng"].h, _lib.ptr(flat), _lib.ptr(st["m"]), _lib.ptr(st["v"]), st["ptrs"], len(st["tensors"]),
                                  lr, self.optimizer.beta1, self.optimizer.beta2, self.optimizer.epsilon, self.global_step + 1,
                                  self.clip_gradient_norm, ctypes.c_void_p(st["ws"].data_ptr() + off), st["need"],
                                  _lib.current_stream()))
        st["store"].version += 1                                # the engine re-packs the updated weights on its next use
        self.global_step += 1

    def __call__(self, data, labels):
        b = self.builder
        b(data, training=True)
        b.create_objectives(labels)
        flat, views = b.backward(allreduce=True)
        self.apply_gradients(flat, views)
        return b.objectives.total_loss()


def create_train_op(builder, optimizer, global_step=0, clip_gradient_norm=0.0):
    """slim.learning.create_train_op(total_loss, optimizer, global_step, clip_gradient_norm=...) (train.py:127-129)."""
    return TrainOp(builder, optimizer, global_step, clip_gradient_norm)
