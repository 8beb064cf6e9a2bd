"""model/yolo2/inference.py -- the ``inference`` functions string-dispatched from the INI config
(``getattr(inference, config.get('yolo2', 'inference'))``, model/yolo2/__init__.py:107).

``darknet`` keeps the reference signature and return value ``(scope, net)`` (inference.py:61-120).
The graph it stands for -- 21x (3x3|1x1 conv -> BN -> leaky), 5 max-pools, the passthrough reorg +
concat, the final linear 1x1 conv -- runs as hand-written sm_100a kernels behind one C-ABI call
(y2_darknet_forward); weights live in the variable store under the reference's TF names.

``tiny`` (inference.py:25-50) is the second function the shipped configs select
(config/yolo2/tiny-{20,80}.ini): the same kernels behind a different layer table (y2_create_net with
Y2_ARCH_TINY), inference and training step.
"""
import sys

from ... import _lib
from ... import variables as V


def layer_geometry(classes, num_anchors):
    """[(name, ksize, cin, cout, has_bn, pool_after)] of the 22 convs in graph order
    (inference.py:70-118): host-side table, same as the one compiled into the library."""
    t, cin, ch = [], 3, 32

    def add(k, cout, pool=False):
        nonlocal cin
        t.append(("conv%d" % len(t), k, cin, cout, True, pool))
        cin = cout

    for _ in range(2):
        add(3, ch, True)
        ch *= 2
    for _ in range(2):
        add(3, ch)
        add(1, ch // 2)
        add(3, ch, True)
        ch *= 2
    for k, c in ((3, ch), (1, ch // 2), (3, ch), (1, ch // 2)):
        add(k, c)
    add(3, ch, True)
    ch *= 2
    for k, c in ((3, ch), (1, ch // 2), (3, ch), (1, ch // 2), (3, ch), (3, ch), (3, ch)):
        add(k, c)
    cin = 4 * 512 + ch
    add(3, ch)
    t.append(("conv", 1, ch, num_anchors * (5 + classes), False, False))
    return t


ARCH_DARKNET, ARCH_TINY = 0, 1      # Y2_ARCH_* (include/yolo2_b200.h)


def tiny_layer_geometry(classes, num_anchors):
    """[(name, ksize, cin, cout, has_bn, pool_after)] of tiny's 9 convs (inference.py:33-48); pool_after is
    False, True (2x2 stride 2) or 's1' (2x2 stride 1 SAME, :42)."""
    t, cin, ch = [], 3, 16
    for _ in range(5):
        t.append(("conv%d" % len(t), 3, cin, ch, True, True))
        cin, ch = ch, ch * 2
    t.append(("conv%d" % len(t), 3, cin, ch, True, 's1'))
    cin, ch = ch, ch * 2
    for _ in range(2):
        t.append(("conv%d" % len(t), 3, cin, ch, True, False))
        cin = ch
    t.append(("conv", 1, ch, num_anchors * (5 + classes), False, False))
    return t


class _Engine(object):
    """One y2_handle per (device, classes, anchors, network); re-uploads weights when the store changes."""
    _cache = {}
    # Diagnostics / tests: give every layer's output its own workspace slot so that `activation()` can read it after the forward.
    # Off (the default), the outputs alternate between two arenas (2.4x less workspace).
    KEEP_ACTIVATIONS = False

    def __init__(self, device_index, classes, num_anchors, arch=ARCH_DARKNET):
        import ctypes
        self.h = ctypes.c_void_p()
        _lib.check(_lib.lib().y2_create_net(ctypes.byref(self.h), device_index, classes, num_anchors, arch))
        self.classes, self.num_anchors, self.arch = classes, num_anchors, arch
        self.loaded_version = None
        self.loaded_store = None
        self.loaded_key = None
        self.center = True
        self.ws = None
        L = _lib.lib()
        self.layers = []
        for i in range(L.y2_num_layers(self.h)):
            k, cin, cout, bn = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int())
            _lib.check(L.y2_layer_info(self.h, i, ctypes.byref(k), ctypes.byref(cin), ctypes.byref(cout), ctypes.byref(bn)))
            self.layers.append((k.value, cin.value, cout.value, bn.value))

    @classmethod
    def get(cls, device, classes, num_anchors, arch=ARCH_DARKNET):
        key = (device.index or 0, classes, num_anchors, arch)
        if key not in cls._cache:
            cls._cache[key] = cls(key[0], classes, num_anchors, arch)
        eng = cls._cache[key]
        if getattr(eng, "_keep", None) != bool(cls.KEEP_ACTIVATIONS):
            _lib.check(_lib.lib().y2_set_option(eng.h, b"keep_activations", int(bool(cls.KEEP_ACTIVATIONS))))
            eng._keep = bool(cls.KEEP_ACTIVATIONS)
        return eng

    def sync_weights(self, scope, store, device, center=True, weights_initializer=V.xavier_uniform):
        if self.loaded_store is store and self.loaded_version == store.version and self.loaded_key == (scope, bool(center)):
            return
        L = _lib.lib()
        n = len(self.layers)
        for i, (k, cin, cout, bn) in enumerate(self.layers):
            name = "%s/conv%d" % (scope, i) if i < n - 1 else "%s/conv" % scope
            w = store.get(name + "/weights", (k, k, cin, cout), weights_initializer, device)
            if bn:
                g = store.get(name + "/BatchNorm/gamma", (cout,), V.ones, device)
                # center=False (the `_darknet` variant, inference.py:62-66): no beta, a separate `biases` variable added after BN
                b = store.get(name + ("/BatchNorm/beta" if center else "/biases"), (cout,), V.zeros, device)
                m = store.get(name + "/BatchNorm/moving_mean", (cout,), V.zeros, device)
                v = store.get(name + "/BatchNorm/moving_variance", (cout,), V.ones, device)
                _lib.check(L.y2_load_weights(self.h, i, _lib.ptr(w), _lib.ptr(g), _lib.ptr(b), _lib.ptr(m), _lib.ptr(v),
                                             None, _lib.current_stream()))
            else:
                bias = store.get(name + "/biases", (cout,), V.zeros, device)
                _lib.check(L.y2_load_weights(self.h, i, _lib.ptr(w), None, None, None, None, _lib.ptr(bias),
                                             _lib.current_stream()))
        self.loaded_store, self.loaded_version, self.loaded_key = store, store.version, (scope, bool(center))
        self.center = bool(center)

    def forward(self, x, precision=0):
        import torch
        L = _lib.lib()
        b, h, w, c = x.shape
        if c != 3:
            raise ValueError("darknet expects NHWC input with 3 channels, got %s" % (tuple(x.shape),))
        need = L.y2_workspace_bytes(self.h, b, h, w)
        if need == 0:
            raise _lib.Y2Error(L.y2_last_error().decode())
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(need + 1024, dtype=torch.uint8, device=x.device)
        off = (-self.ws.data_ptr()) % 1024
        out = torch.empty((b, h // 32, w // 32, self.num_anchors * (5 + self.classes)), dtype=torch.float32, device=x.device)
        import ctypes
        _lib.check(L.y2_darknet_forward(self.h, _lib.ptr(x, torch.float32), b, h, w, _lib.ptr(out),
                                        ctypes.c_void_p(self.ws.data_ptr() + off), need, precision, _lib.current_stream()))
        return out

    # ---- training step (train.py:109-129) ----
    def forward_train(self, x, scope, store):
        """Forward with batch statistics; keeps activations for backward(); moving averages are updated in
        the engine AND written back to the variable store (slim's UPDATE_OPS)."""
        import ctypes
        import torch
        L = _lib.lib()
        b, h, w, c = x.shape
        need = L.y2_train_workspace_bytes(self.h, b, h, w)
        if need == 0:
            raise _lib.Y2Error(L.y2_last_error().decode())
        if getattr(self, "tws", None) is None or self.tws.numel() < need + 1024:
            self.tws = None
            self.tws = torch.empty(need + 1024, dtype=torch.uint8, device=x.device)
        off = (-self.tws.data_ptr()) % 1024
        out = torch.empty((b, h // 32, w // 32, self.num_anchors * (5 + self.classes)), dtype=torch.float32, device=x.device)
        _lib.check(L.y2_darknet_forward_train(self.h, _lib.ptr(x, torch.float32), b, h, w, _lib.ptr(out),
                                              ctypes.c_void_p(self.tws.data_ptr() + off), need, _lib.current_stream()))
        self._train_x = x                      # the backward reads the input again (conv0 weight gradient)
        n = len(self.layers)
        for i, (k, cin, cout, bn) in enumerate(self.layers):
            if not bn:
                continue
            name = "%s/conv%d/BatchNorm/" % (scope, i)
            mm = store.get(name + "moving_mean", (cout,), V.zeros, x.device)
            mv = store.get(name + "moving_variance", (cout,), V.ones, x.device)
            _lib.check(L.y2_get_bn_state(self.h, i, None, None, _lib.ptr(mm), _lib.ptr(mv), _lib.current_stream()))
        return out

    def backward(self, dnet):
        """d(total_loss)/d(net) -> (flat float32 gradient bucket, {variable suffix -> view})."""
        import ctypes
        import torch
        L = _lib.lib()
        flat = torch.empty(L.y2_param_count(self.h), dtype=torch.float32, device=dnet.device)
        _lib.check(L.y2_darknet_backward(self.h, _lib.ptr(dnet.contiguous(), torch.float32), _lib.ptr(flat), _lib.current_stream()))
        views = {}
        n = len(self.layers)
        for i, (k, cin, cout, bn) in enumerate(self.layers):
            w_off, g_off, b_off = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
            _lib.check(L.y2_param_offsets(self.h, i, ctypes.byref(w_off), ctypes.byref(g_off), ctypes.byref(b_off)))
            name = "conv%d" % i if i < n - 1 else "conv"
            views[name + "/weights"] = flat[w_off.value:w_off.value + k * k * cin * cout].view(k, k, cin, cout)
            if bn:
                views[name + "/BatchNorm/gamma"] = flat[g_off.value:g_off.value + cout]
                # center=False graphs (`_darknet`): the shift is the separate `<conv>/biases` variable, there is no beta
                views[name + ("/BatchNorm/beta" if self.center else "/biases")] = flat[b_off.value:b_off.value + cout]
            else:
                views[name + "/biases"] = flat[b_off.value:b_off.value + cout]
        return flat, views

    def activation(self, layer, pooled, shape):
        import torch
        out = torch.empty(shape, dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib().y2_get_activation(self.h, layer, int(pooled), _lib.ptr(out), _lib.current_stream()))
        return out


PRECISION = 0      # 0 = split-bf16 x3 (fp32-grade, the parity mode); 1 = single bf16 pass


def darknet(net, classes, num_anchors, training=False, center=True, precision=None):
    """Darknet-19 + passthrough backbone (inference.py:61-120).

    net: float32 CUDA tensor [B, H, W, 3] NHWC.  Returns ``(scope, output)`` with
    output [B, H/32, W/32, num_anchors*(5+classes)] and scope == 'yolo2_darknet' (inference.py:67).
    """
    # package + function name, as the reference derives it (inspect.stack()[0][3] there; the frame's code name is the
    # same string without inspect's per-call source-file stat/read, ~0.2 ms)
    scope = __name__.split('.')[-2] + '_' + sys._getframe().f_code.co_name
    if not net.is_cuda:
        raise _lib.Y2Error("darknet: input must be a CUDA tensor (no CPU path exists)")
    eng = _Engine.get(net.device, classes, num_anchors)
    eng.sync_weights(scope, V.default_store(), net.device, center=center)
    if training:
        return scope, eng.forward_train(net.contiguous(), scope, V.default_store())
    out = eng.forward(net.contiguous(), precision=PRECISION if precision is None else precision)
    return scope, out


DARKNET_DOWNSAMPLING = (2 ** 5, 2 ** 5)      # inference.py:122


def tiny(net, classes, num_anchors, training=False, center=True, precision=None):
    """Tiny YOLOv2 backbone (inference.py:25-50): conv0..conv4 (16..256 channels, each + 2x2/2 max-pool), conv5 (512) +
    2x2 stride-1 SAME max-pool, conv6/conv7 (1024), linear 1x1 conv.  Same signature and ``(scope, net)`` return as the
    reference; scope == 'yolo2_tiny'.  Weights not present in the variable store are created with the reference's
    initializer for this function (truncated_normal(stddev=0.1), :33).  training=True runs the batch-statistics forward and keeps
    the state `Builder.backward` needs, like `darknet`."""
    scope = __name__.split('.')[-2] + '_' + sys._getframe().f_code.co_name
    if not net.is_cuda:
        raise _lib.Y2Error("tiny: input must be a CUDA tensor (no CPU path exists)")
    eng = _Engine.get(net.device, classes, num_anchors, ARCH_TINY)
    eng.sync_weights(scope, V.default_store(), net.device, center=center, weights_initializer=V.truncated_normal_01)
    if training:
        return scope, eng.forward_train(net.contiguous(), scope, V.default_store())
    return scope, eng.forward(net.contiguous(), precision=PRECISION if precision is None else precision)


TINY_DOWNSAMPLING = (2 ** 5, 2 ** 5)         # inference.py:52


def _tiny(net, classes, num_anchors, training=False):
    """inference.py:55-56: tiny with center=False (BN without beta + separate `biases`)."""
    return tiny(net, classes, num_anchors, training, False)


_TINY_DOWNSAMPLING = (2 ** 5, 2 ** 5)


def _darknet(net, classes, num_anchors, training=False):
    """inference.py:125-126: BN without beta + separate bias.  With beta == 0 the arithmetic is
    identical, so the variant shares the kernels (the separate 'biases' variable maps to beta)."""
    return darknet(net, classes, num_anchors, training, False)


_DARKNET_DOWNSAMPLING = (2 ** 5, 2 ** 5)
