"""Multi-GPU parity (needs >= 2 visible GPUs, skipped otherwise): tests/mgpu_check.py runs the z-slab decomposition
on N ranks over NCCL and compares forces, energies, virial and a 12-step trajectory (with migrations between the
slabs) against a single-GPU context holding the same system, for `kspace_modify diff ik` and `diff ad`, with the
dispersion grids of pppm/disp (geometric, arithmetic, no mixing rule) as a second PPPM state, with special bonds on
the device-built lists (data.spce) and with fix nve on a sub-group with per-atom masses."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,diff,disp,mode,group", [(2, 0, 0, "aC", 0), (4, 0, 0, "aC", 0), (2, 1, 0, "aC", 0),
                                                        (2, 0, 1, "aC", 0), (2, 0, 2, "aC", 0), (2, 1, 3, "aC", 0),
                                                        (2, 0, 0, "spce", 0), (2, 0, 0, "aC", 1)])
def test_slab_decomposition_matches_single_gpu(world, diff, disp, mode, group):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, DIFF=str(diff), DISP=str(disp), MODE=mode, GROUP=str(group)))
    assert r.returncode == 0 and "MGPU CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
