"""Image pre- / post-processing kernels on the device — exact parity with the oracle, the golden
vectors of Pillow / torchvision and the libraries themselves, through the C ABI (`mb_image_preprocess_u8`,
`mb_image_postprocess_u8`) and the drop-in classes (`mingtok.utils.CenterCropProcessor`).

The checks run in tests/native/preprocess_gpu_worker.py, a process of their own (first hardware run: the driver's
round-1 GPU suite, green); the same device code is also walked on the CPU through the header it shares with the emulation
(tests/test_preprocess_cpu.py).
"""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_preprocess_on_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "preprocess_gpu_worker.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "preprocess worker ok" in r.stdout, (r.stdout[-2000:] + "\n" + r.stderr[-3000:])
