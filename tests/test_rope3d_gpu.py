"""3-D multimodal RoPE kernel on the device (the config-gated `rope_scaling.type == "3D"` variant):
`mb_rope3d_kv_append` against the oracle / the reference's golden vectors, and the gated path through
`BailingMoeModel.forward_tokens`.  Runs in tests/native/rope3d_gpu_worker.py, a process of its own (first hardware run:
the driver's round-1 GPU suite, green); the arithmetic, cache layout and section selection are also verified on the CPU
through the header shared with the emulation (tests/test_rope3d_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_rope3d_on_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "rope3d_gpu_worker.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "rope3d worker ok" in r.stdout, (r.stdout[-2000:] + "\n" + r.stderr[-3000:])
