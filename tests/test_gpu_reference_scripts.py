"""The reference's own train_spatial_query.py / test_spatial_query.py run UNMODIFIED against this repository's modules
(tools/run_ref_script.py: only the process environment is adapted).  Needs the reference tree on the
box (baseline/_ref); skipped without it."""
import os
import subprocess
import sys

import pytest
import torch

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
LAUNCH = os.path.join(ROOT, "tools", "run_ref_script.py")


def _ref():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        from run_ref_script import find_reference
    finally:
        sys.path.pop(0)
    return find_reference()


def _run(cmd, timeout=900):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    return r.stdout + r.stderr


@pytest.fixture(scope="module")
def trained(tmp_path_factory):
    if _ref() is None:
        pytest.skip("no reference tree on this machine (baseline/_ref)")
    wd = str(tmp_path_factory.mktemp("refrun"))
    out = _run([sys.executable, LAUNCH, "--workdir", wd, "train_spatial_query.py", "synthetic-lmdb", "--iter", "1",
                "--batch", "4", "--size", "64", "--n_sample", "4", "--exp_name", "unmod"])
    return wd, out


def test_train_script_runs_unmodified(trained):
    wd, out = trained
    ckpt = os.path.join(wd, "out", "unmod", "checkpoint", "000000.pt")
    assert os.path.isfile(ckpt), out[-2000:]
    assert os.path.isfile(os.path.join(wd, "out", "unmod", "sample", "000000.png"))
    sd = torch.load(ckpt, map_location="cpu")
    assert sorted(sd) == ["d", "d_optim", "g", "g_ema", "g_optim"]
    assert all(torch.isfinite(v).all() for v in sd["g"].values() if v.is_floating_point())
    assert "Generator params count: " in out


def test_inference_script_runs_unmodified(trained):
    wd, _ = trained
    ckpt = os.path.join(wd, "out", "unmod", "checkpoint", "000000.pt")
    _run([sys.executable, LAUNCH, "--workdir", wd, "test_spatial_query.py", "--ckpt", ckpt, "--size", "64", "--n_sample", "2",
          "--loop_num", "2", "--sample", "--swap_z", "--swap_p", "--interp", "--interp_num", "6"])
    vis = os.path.join(wd, "generation", "visual", "unmod", "0")
    files = os.listdir(vis)
    assert "swap_z.png" in files and "0.png" in files and len(files) >= 4, files


def test_train_script_under_stock_ddp_two_ranks(tmp_path):
    """torchrun --nproc-per-node 2: the script wraps our modules in torch DistributedDataParallel
    (find_unused_parameters=True, train_spatial_query.py:495-509) and all-reduces the losses."""
    if _ref() is None:
        pytest.skip("no reference tree on this machine (baseline/_ref)")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                "127.0.0.1", "--master-port", "29541", LAUNCH, "--workdir", str(tmp_path), "train_spatial_query.py",
                "synthetic-lmdb", "--iter", "1", "--batch", "4", "--size", "64", "--n_sample", "4", "--exp_name", "ddp"])
    assert os.path.isfile(os.path.join(str(tmp_path), "out", "ddp", "checkpoint", "000000.pt")), out[-2000:]
