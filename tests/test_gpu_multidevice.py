"""mvf_gpu_init_multi: one process driving several B200s (tests/multidevice_check.py when the box has >= 2 GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import metrovector_b200 as mv

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_init_multi_with_one_device_is_a_plain_engine():
    eng = mv.GpuEngine([0])
    assert eng.device_count() == 1
    db = mv.synth_fill_host(3, 5, 0, 3000, 64)
    gs = eng.from_host(db, mv.DataType.UInt8, mv.DistanceMetric.L2)
    ids, _, isc = gs.search(db[7], 3)
    assert ids[0][0] == 7 and isc[0][0] == 0
    gs.free()
    eng.shutdown()
    with pytest.raises(mv.Build):
        mv.GpuEngine([0, 0])
    with pytest.raises(mv.Build):
        mv.GpuEngine([])


def test_two_engines_in_one_process_on_different_devices():
    # the shared-memory opt-in of the merge kernels is a per-device attribute (it used to be latched per process)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = []
    for dev in (0, 1):
        eng = mv.GpuEngine(dev)
        db = mv.synth_fill_host(0, 5, 0, 4000, 768)
        gs = eng.from_host(db, mv.DataType.Float32, mv.DistanceMetric.L2)
        res.append(gs.search(db[:9], 100))          # 9 queries x k=100: the tournament merge with > 48 KB of shared memory
        gs.free()
        eng.shutdown()
    assert np.array_equal(res[0][0], res[1][0])


def test_group_engine_matches_single_gpu_and_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (profiles/r2_multidevice_*.txt hold the 2- and 8-GPU runs)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multidevice_check.py")], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multidevice ok" in out.stdout
