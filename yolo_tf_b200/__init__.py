"""yolo_tf_b200 -- B200-native YOLOv2 / Darknet-19 detection hot path.

A from-scratch sm_100a implementation (tcgen05 + TMA implicit-GEMM convs, fused head / loss /
NMS kernels in ``csrc/``) behind the Python call surface of ruiminshen/yolo-tf that
``train.py`` / ``detect.py`` drive:

    model.yolo2.inference.darknet     model/yolo2/inference.py:61-120
    model.yolo2.Builder/Model/Objectives   model/yolo2/__init__.py:28-119
    utils.postprocess.non_max_suppress     utils/postprocess.py:39-51

PyTorch tensors are used only as device-memory containers; all arithmetic runs in
``csrc/libyolo2_b200.so`` through the C ABI declared in ``include/yolo2_b200.h``.
There is no CPU fallback: importing works anywhere, calling requires a B200 and the built library.
"""
__version__ = "0.1.0"
