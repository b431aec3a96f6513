"""GPU, two PROCESSES on two GPUs: the fused partition over real ranks — CUDA IPC mappings of the other rank's images, peer stores over NVLink,
cuStreamWriteValue32 / cuStreamWaitValue32 flag words between processes — against the same frames on one GPU, bit for bit (raw masks,
reflections, denoised). tests/test_partition_gpu.py runs the same kernels with the ranks as contexts of one process; this one covers the
inter-process plumbing. Skipped on a box with a single GPU (the driver's single-GPU test tier); runs under `gpurun --gpus 2`."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return n.value if cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (cuDeviceGetCount < 2)")
@pytest.mark.parametrize("world", [2])
def test_fused_partition_over_two_processes_is_bit_exact(world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "fused_partition_parity.py"), "960", "544", "60000", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PARITY OK (bit-exact)" in r.stdout
