"""N-GPU domain decomposition == 1 GPU, bit for bit (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config,n,cuts,transport", [
    ("default_slab", 1_000_000, "equal", "nccl"),
    ("optically_thick", 20_000, "equal", "nccl"),
    ("default_slab", 1_000_000, "uneven", "nccl"),
    ("default_slab", 1_000_000, "uneven", "p2p"),      # kernel stores into the peer's inbox
    ("optically_thick", 20_000, "equal", "p2p"),
])
def test_multi_gpu_equals_single_gpu(gpu, mcb_lib, config, n, cuts, transport):
    ngpu = mcb_lib.mcb200_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    K = min(ngpu, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={K}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "check_multi_gpu.py"), str(n), str(1 << 17), config, cuts, transport]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "tally_bit_exact=True" in res.stdout
    assert f"transport={transport}" in res.stdout    # no silent fall-back to the other transport


@pytest.mark.parametrize("case,cuts,windows", [
    ("default_slab_2e6", "equal", 0),
    ("default_slab_2e6", "uneven", 0),
    ("default_slab_1e5", "equal", 2),
    ("absorption_dominated_2e5", "equal", 0),
    ("optically_thick_2e4", "uneven", 0),
])
def test_world_run_equals_oracle_digest(gpu, mcb_lib, case, cuts, windows):
    """the persistent world kernels (one per GPU, rings over NVLink, device-side termination)
    against the oracle digest committed in tests/golden/world_digest.json"""
    ngpu = mcb_lib.mcb200_device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    K = min(ngpu, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={K}",
           "--master-addr", "127.0.0.1", "--master-port", "29534",
           os.path.join(ROOT, "tools", "check_world.py"), case, cuts, str(windows)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "[world parity] ok=True" in res.stdout
