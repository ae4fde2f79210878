"""-m gpu, needs >= 2 GPUs (skipped on a one-GPU box): the multi-GPU parity checkers under torchrun — the Ulysses forward
(Wan, CogVideoX; peer and NCCL exchange) against the un-sharded model, and the row-parallel VAE against the un-sharded
VAE (bit-identical)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, *args, port=29541):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.fixture(scope="module")
def nproc():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return 2 if n < 4 else (4 if n < 8 else 8)


def test_vae_row_parallel_bit_identical(nproc):
    r = _torchrun(nproc, "vae_sp_check.py")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("VAE_SP ")][-1]
    import json

    res = json.loads(line[len("VAE_SP "):])
    assert res["ok"] and all(c["equal"] and c["encode_max_abs_diff_vs_unsharded"] == 0.0 for c in res["cases"])


@pytest.mark.parametrize("which", ["wan", "cog"])
def test_ulysses_forward_matches_unsharded(nproc, which):
    r = _torchrun(nproc, "sp_check.py", *(["cog"] if which == "cog" else []), port=29542)
    assert r.returncode == 0 and "SP_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
