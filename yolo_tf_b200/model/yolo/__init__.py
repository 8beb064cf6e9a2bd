"""Subset of the reference's model/yolo package that the YOLOv2 path touches."""
import numpy as np


def calc_cell_xy(cell_height, cell_width, dtype=np.float32):
    """[H, W, 2] grid holding (x, y) of every cell -- model/yolo/__init__.py:29-34.
    (The decode kernel derives the same values from the cell index; this host helper keeps the
    reference's public function.)"""
    grid = np.empty([cell_height, cell_width, 2], dtype=dtype)
    grid[:, :, 0] = np.arange(cell_width, dtype=dtype).reshape(1, cell_width)
    grid[:, :, 1] = np.arange(cell_height, dtype=dtype).reshape(cell_height, 1)
    return grid
