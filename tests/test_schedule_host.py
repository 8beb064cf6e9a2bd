"""CPU (-m "not gpu"): the tile schedule the tcgen05 kernels walk -- choose_schedule (hybrid data-parallel + stream-K cost
model, csrc/y2_internal.h) and SegIter / CapIter (csrc/y2_ptx.cuh: stream-K ranges cut into accumulation chains of at
most `kcap` k-blocks) -- compiled for the HOST from the same headers the kernels include (tests/host/schedule_harness.cu)
and checked exhaustively: exact cover of every (tile, k-block), chain length <= cap, contiguity, and the two ordering
invariants the hand-off relies on (partials are published first, heads with contributors come last)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schedule_and_chain_cap_cover_every_kblock_once(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "schedule_harness")
    src = os.path.join(ROOT, "tests", "host", "schedule_harness.cu")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "SCHEDULE CHECK OK" in out.stdout
    # the plans of the bench configuration, as documented in DESIGN 4.1 / 4.8
    assert "tiles=172 KB=432 workers=148 cap=32 -> dp_tiles=148 sk_workers=148" in out.stdout     # conv20: hybrid
    assert "tiles=86 KB=72 workers=74 cap=32 -> dp_tiles=0 sk_workers=74" in out.stdout            # conv13 as CTA pairs: stream-K
    assert "tiles=338 KB=18 workers=74 cap=32 -> dp_tiles=338 sk_workers=0" in out.stdout          # conv5 as CTA pairs: data-parallel
