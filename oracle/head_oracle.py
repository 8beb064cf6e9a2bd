"""CPU oracle for the YOLOv2 head: decode, objectives (loss) and its gradient.
TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs).

Restates with numpy (float32 by default, float64 twin via ``dtype``):

* ``decode_oracle``      <- ``Model.__init__``       model/yolo2/__init__.py:28-59
                            + ``calc_cell_xy``        model/yolo/__init__.py:29-34
* ``objectives_oracle``  <- ``Objectives.__init__``  model/yolo2/__init__.py:62-94
* ``total_loss_oracle``  <- ``Builder.create_objectives`` weighting :114-119 with
                            ``[yolo2_hparam]`` config.ini:98-102
* ``loss_grad_oracle``   closed-form d(total)/d(inputs);
  ``loss_grad_autograd`` is the torch-autograd twin used to validate it.
* ``transform_labels_oracle`` <- ``transform_labels`` utils/data/__init__.py:112-145
  (label layout the loss consumes; ``np.int`` replaced by ``int``).

PINNED TO THE REFERENCE'S SOURCE, NOT TO TENSORFLOW'S ARITHMETIC: tests/golden/head_reference.npz holds
the outputs of the reference's own ``Model`` / ``Objectives`` classes (and autograd through them),
compiled from the reference file and run with a torch stand-in for the ~20 TF ops they call
(tests/golden/make_head_golden.py); decode, objectives and the closed-form gradient match them to
1e-12 / 1e-10 in float64 (tests/test_head_reference_golden.py).  TensorFlow 1.0 itself is not
installable here, so its kernel arithmetic (summation order, exp / sigmoid ulps) stays PARITY
UNPINNED.  TF semantics assumed: softmax subtracts the row max; ``tf.equal`` masks carry no
gradient; reductions are float32 sums.
"""
import numpy as np

HPARAM_DEFAULT = {"prob": 1.0, "iou_best": 5.0, "iou_normal": 1.0, "coords": 1.0}   # config.ini:98-102
ANCHORS_COCO = np.array([[0.738768, 0.874946], [2.42204, 2.65704], [4.30971, 7.04493],
                         [10.246, 4.59428], [12.6868, 11.8741]])                      # config/yolo2/anchors/coco.tsv
ANCHORS_VOC = np.array([[1.08, 1.19], [3.42, 4.41], [6.63, 11.38], [9.42, 5.11], [16.62, 10.52]])  # .../voc.tsv


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def cell_xy_grid(cell_height, cell_width, dtype=np.float32):
    """model/yolo/__init__.py:29-34: [H,W,2] with (x, y) per cell."""
    g = np.zeros([cell_height, cell_width, 2], dtype=dtype)
    g[..., 0] = np.arange(cell_width, dtype=dtype)[None, :]
    g[..., 1] = np.arange(cell_height, dtype=dtype)[:, None]
    return g


def decode_oracle(net, classes, anchors, training=False, dtype=np.float32):
    """net [B,Hc,Wc,A*(5+C)] -> dict of the Model attributes (model/yolo2/__init__.py:28-59)."""
    net = np.asarray(net, dtype=dtype)
    anchors = np.asarray(anchors, dtype=dtype)
    b, hc, wc, _ = net.shape
    cells, a = hc * wc, len(anchors)
    inputs = net.reshape(b, cells, a, 5 + classes)                         # :32
    sig = _sigmoid(inputs[..., :3]).astype(dtype)                          # :36
    m = {"cell_height": hc, "cell_width": wc}
    m["iou"] = sig[..., 0]                                                 # :37
    m["offset_xy"] = sig[..., 1:3]                                         # :38
    m["wh"] = (np.exp(inputs[..., 3:5]) * anchors.reshape(1, 1, a, 2)).astype(dtype)   # :40
    z = inputs[..., 5:]
    e = np.exp(z - z.max(-1, keepdims=True))
    m["prob"] = (e / e.sum(-1, keepdims=True)).astype(dtype)               # :42
    m["areas"] = m["wh"][..., 0] * m["wh"][..., 1]                         # :43
    half = m["wh"] / dtype(2)                                              # :44
    m["offset_xy_min"] = m["offset_xy"] - half                             # :45
    m["offset_xy_max"] = m["offset_xy"] + half                             # :46
    m["wh01"] = m["wh"] / np.array([wc, hc], dtype=dtype).reshape(1, 1, 1, 2)   # :47
    m["wh01_sqrt"] = np.sqrt(m["wh01"])                                    # :48
    m["coords"] = np.concatenate([m["offset_xy"], m["wh01_sqrt"]], -1)     # :49
    if not training:                                                       # :50-56
        cxy = cell_xy_grid(hc, wc, dtype).reshape(1, cells, 1, 2)
        m["xy"] = cxy + m["offset_xy"]
        m["xy_min"] = cxy + m["offset_xy_min"]
        m["xy_max"] = cxy + m["offset_xy_max"]
        m["conf"] = m["iou"][..., None] * m["prob"]
    return m


def objectives_oracle(m, labels, dtype=np.float32):
    """model/yolo2/__init__.py:62-94.  labels = (mask[B,cells,1], prob[B,cells,1,C],
    coords[B,cells,1,4], offset_xy_min[B,cells,1,2], offset_xy_max[B,cells,1,2], areas[B,cells,1]).
    Returns (objectives dict, aux dict with iou / mask_best)."""
    mask, prob, coords, oxy_min, oxy_max, areas = [np.asarray(t, dtype=dtype) for t in labels]
    lo = np.maximum(m["offset_xy_min"], oxy_min)                           # :73
    hi = np.minimum(m["offset_xy_max"], oxy_max)                           # :74
    wh = np.maximum(hi - lo, dtype(0))                                     # :75
    inter = wh[..., 0] * wh[..., 1]                                        # :76
    union = np.maximum(areas + m["areas"] - inter, dtype(1e-10))           # :77
    iou = inter / union                                                    # :78
    best = (iou == iou.max(2, keepdims=True)).astype(dtype)                # :80-81
    mask_best = mask * best                                                # :82
    mask_normal = dtype(1) - mask_best                                     # :83
    iou_dist = (m["iou"] - mask_best) ** 2                                 # :85
    coords_dist = (m["coords"] - coords) ** 2                              # :86
    prob_dist = (m["prob"] - prob) ** 2                                    # :87
    cnt = dtype(np.multiply.reduce(iou_dist.shape))                        # :89
    obj = {
        "iou_best": (mask_best * iou_dist).sum(dtype=dtype) / cnt,          # :90
        "iou_normal": (mask_normal * iou_dist).sum(dtype=dtype) / cnt,      # :91
        "coords": (mask_best[..., None] * coords_dist).sum(dtype=dtype) / cnt,   # :93
        "prob": (mask_best[..., None] * prob_dist).sum(dtype=dtype) / cnt,       # :94
    }
    return obj, {"iou": iou, "mask_best": mask_best, "mask_normal": mask_normal, "cnt": cnt}


def total_loss_oracle(obj, hparam=HPARAM_DEFAULT):
    return sum(obj[k] * hparam[k] for k in ("prob", "iou_best", "iou_normal", "coords"))


def loss_grad_oracle(net, classes, anchors, labels, hparam=HPARAM_DEFAULT, dtype=np.float32):
    """Closed-form d(total_loss)/d(net), shape of net.  Returns (objectives, grad)."""
    m = decode_oracle(net, classes, anchors, training=True, dtype=dtype)
    obj, aux = objectives_oracle(m, labels, dtype=dtype)
    mask, prob, coords, _, _, _ = [np.asarray(t, dtype=dtype) for t in labels]
    mb, mn, cnt = aux["mask_best"], aux["mask_normal"], aux["cnt"]
    b, cells, a = mb.shape
    g = np.zeros((b, cells, a, 5 + classes), dtype=dtype)
    s0 = m["iou"]
    w_o = (dtype(hparam["iou_best"]) * mb + dtype(hparam["iou_normal"]) * mn) / cnt
    g[..., 0] = dtype(2) * (s0 - mb) * s0 * (dtype(1) - s0) * w_o
    sxy = m["offset_xy"]
    g[..., 1:3] = dtype(hparam["coords"]) * mb[..., None] * dtype(2) * (sxy - coords[..., 0:2]) * sxy * (dtype(1) - sxy) / cnt
    s = m["wh01_sqrt"]
    g[..., 3:5] = dtype(hparam["coords"]) * mb[..., None] * dtype(2) * (s - coords[..., 2:4]) * dtype(0.5) * s / cnt
    p = m["prob"]
    q = dtype(hparam["prob"]) * mb[..., None] * dtype(2) * (p - prob) / cnt
    g[..., 5:] = p * (q - (q * p).sum(-1, keepdims=True))
    return obj, g.reshape(np.asarray(net).shape)


def loss_grad_autograd(net, classes, anchors, labels, hparam=HPARAM_DEFAULT):
    """float64 torch-autograd twin of the objectives, for validating the closed form."""
    import torch
    x = torch.tensor(np.asarray(net, dtype=np.float64), requires_grad=True)
    anc = torch.tensor(np.asarray(anchors, dtype=np.float64))
    b, hc, wc, _ = x.shape
    cells, a = hc * wc, anc.shape[0]
    inp = x.reshape(b, cells, a, 5 + classes)
    sig = torch.sigmoid(inp[..., :3])
    iou_p, oxy = sig[..., 0], sig[..., 1:3]
    wh = torch.exp(inp[..., 3:5]) * anc.reshape(1, 1, a, 2)
    prob_p = torch.softmax(inp[..., 5:], -1)
    areas_p = wh[..., 0] * wh[..., 1]
    omin, omax = oxy - wh / 2, oxy + wh / 2
    wh01s = torch.sqrt(wh / torch.tensor([wc, hc], dtype=torch.float64).reshape(1, 1, 1, 2))
    coords_p = torch.cat([oxy, wh01s], -1)
    mask, prob, coords, tmin, tmax, areas = [torch.tensor(np.asarray(t, dtype=np.float64)) for t in labels]
    lo, hi = torch.maximum(omin, tmin), torch.minimum(omax, tmax)
    iwh = torch.clamp(hi - lo, min=0.0)
    inter = iwh[..., 0] * iwh[..., 1]
    iou = inter / torch.clamp(areas + areas_p - inter, min=1e-10)
    best = (iou == iou.max(2, keepdim=True).values).double().detach()
    mb = mask * best
    mn = 1 - mb
    cnt = float(b * cells * a)
    obj = {
        "iou_best": (mb * (iou_p - mb) ** 2).sum() / cnt,
        "iou_normal": (mn * (iou_p - mb) ** 2).sum() / cnt,
        "coords": (mb[..., None] * (coords_p - coords) ** 2).sum() / cnt,
        "prob": (mb[..., None] * (prob_p - prob) ** 2).sum() / cnt,
    }
    total = sum(obj[k] * hparam[k] for k in obj)
    total.backward()
    return {k: float(v.detach()) for k, v in obj.items()}, x.grad.numpy()


def transform_labels_oracle(objects_class, objects_coord, classes, cell_width, cell_height, dtype=np.float32):
    """utils/data/__init__.py:112-145 (one image).  objects_coord [n,4] = (xmin,ymin,xmax,ymax) in [0,1]."""
    cells = cell_height * cell_width
    mask = np.zeros([cells, 1], dtype=dtype)
    prob = np.zeros([cells, 1, classes], dtype=dtype)
    coords = np.zeros([cells, 1, 4], dtype=dtype)
    oxy_min = np.zeros([cells, 1, 2], dtype=dtype)
    oxy_max = np.zeros([cells, 1, 2], dtype=dtype)
    objects_coord = np.asarray(objects_coord)
    objects_class = np.asarray(objects_class)
    assert len(objects_class) == len(objects_coord)
    xmin, ymin, xmax, ymax = objects_coord.T
    x = cell_width * (xmin + xmax) / 2
    y = cell_height * (ymin + ymax) / 2
    ix, iy = np.floor(x), np.floor(y)
    off_x, off_y = x - ix, y - iy
    w, h = xmax - xmin, ymax - ymin
    index = (iy * cell_width + ix).astype(int)
    mask[index, :] = 1
    prob[index, :, objects_class] = 1
    coords[index, 0, 0] = off_x
    coords[index, 0, 1] = off_y
    coords[index, 0, 2] = np.sqrt(w)
    coords[index, 0, 3] = np.sqrt(h)
    _w = w / 2 * cell_width
    _h = h / 2 * cell_height
    oxy_min[index, 0, 0] = off_x - _w
    oxy_min[index, 0, 1] = off_y - _h
    oxy_max[index, 0, 0] = off_x + _w
    oxy_max[index, 0, 1] = off_y + _h
    wh = oxy_max - oxy_min
    assert np.all(wh >= 0)
    areas = np.multiply.reduce(wh, -1)
    return mask, prob, coords, oxy_min, oxy_max, areas


def synthetic_labels(batch, classes, cell_width, cell_height, seed=3):
    """n~U{1..8} objects/img, class U{0..C-1}, centre U(0,1)^2,
    w,h~U(0.05,0.6) clipped to the image, encoded with transform_labels_oracle. Returns the 6 batched tensors."""
    rs = np.random.RandomState(seed)
    outs = [[] for _ in range(6)]
    for _ in range(batch):
        n = rs.randint(1, 9)
        cls = rs.randint(0, classes, size=n)
        cx, cy = rs.uniform(0, 1, size=n), rs.uniform(0, 1, size=n)
        w, h = rs.uniform(0.05, 0.6, size=n), rs.uniform(0.05, 0.6, size=n)
        xmin, xmax = np.clip(cx - w / 2, 0, 1 - 1e-6), np.clip(cx + w / 2, 0, 1 - 1e-6)
        ymin, ymax = np.clip(cy - h / 2, 0, 1 - 1e-6), np.clip(cy + h / 2, 0, 1 - 1e-6)
        lab = transform_labels_oracle(cls, np.stack([xmin, ymin, xmax, ymax], 1), classes, cell_width, cell_height)
        for o, t in zip(outs, lab):
            o.append(t)
    return tuple(np.stack(o, 0) for o in outs)
