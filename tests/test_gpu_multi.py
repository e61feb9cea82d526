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
