"""GPU (-m gpu): kernel-level parity of the C ABI entry points against plain PyTorch fp32 references
(tools/gpu_probe.py holds the cases: tcgen05 conv fwd/dgrad/stride-2 incl. the halo / swap / CTA-pair modes, wgrad,
GroupNorm/SiLU, layout, pooling, and the row-shifted UMMA descriptor property the halo-tile conv relies on)."""
import importlib
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    return importlib.import_module("gpu_probe")


@pytest.mark.parametrize("group", ["gemm", "conv", "conv2", "wgrad", "elem", "lpips", "up", "stats", "fat", "shift"])
def test_kernel_group(probe, group):
    if group == "shift" and os.environ.get("VQB_DEBUG_LIB", "0") != "1":
        pytest.skip("the bring-up kernel lives in libvqb200_dbg.so: covered by test_debug_library_groups")
    assert getattr(probe, "group_" + group)(), f"kernel parity group {group} has failures (see stdout)"


@pytest.mark.parametrize("group", ["shift", "conv"])
def test_debug_library_groups(group):
    """The -DVQB_DEBUG build (libvqb200_dbg.so): the row-shifted UMMA descriptor property the halo-tile conv relies on
    (csrc/dbg_shift.cu) and the experimental swap / CTA-pair conv modes, in a subprocess that loads that library."""
    import subprocess

    lib = os.path.join(ROOT, "vqgan-training_b200", "libvqb200_dbg.so")
    if not os.path.exists(lib):
        pytest.skip("libvqb200_dbg.so not built (python vqgan-training_b200/build_native.py --debug)")
    env = dict(os.environ, VQB_DEBUG_LIB="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_probe.py"), group], env=env,
                       capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-2000:]
