"""The reference's own trainer drives the drop-in modules unchanged (BASELINE.json north_star: "so train.py / embedding.py
call them unchanged"; SURVEY.md Appendix C, last row).

Needs the reference checkout (/root/reference: present in the build container, absent on the GPU box → skipped there).  Two
subprocesses run the unmodified reference pipeline on UCI with identical seeds — once with the reference's own layers/models,
once after ctgcn_b200.install_as_reference_modules() — and the exported embeddings and saved checkpoints must agree.
CPU only: the CUDA entry points are replaced by the oracle-backed stand-in (tests/fake_backend.py); what is under test here is
everything between the reference's trainer and the C-ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _run(mode, outdir, method, T, epochs):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "dropin_trainer_check.py"), mode, str(outdir), method, str(T), str(epochs)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "dropin_trainer_check OK" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "data", "uci")), reason="needs the reference checkout with its UCI data")
@pytest.mark.parametrize("method,T", [("CTGCN-C", 2), ("CTGCN-S", 2), ("CGCN-C", 1)])
def test_reference_trainer_with_swapped_modules(method, T, tmp_path, lib):
    pd = pytest.importorskip("pandas")
    _run("ref", tmp_path / "ref", method, T, 2)
    _run("dropin", tmp_path / "dropin", method, T, 2)
    files = sorted(f for f in os.listdir(tmp_path / "ref") if f.endswith(".csv"))
    assert len(files) == T and files == sorted(f for f in os.listdir(tmp_path / "dropin") if f.endswith(".csv"))
    for f in files:
        a = pd.read_csv(tmp_path / "ref" / f, sep="\t", index_col=0)
        b = pd.read_csv(tmp_path / "dropin" / f, sep="\t", index_col=0)
        assert list(a.index) == list(b.index) and a.shape == b.shape == (1899, 128)
        err = np.linalg.norm(a.values - b.values) / np.linalg.norm(a.values)
        assert err < 1e-4, (f, err)
    sa, sb = np.load(tmp_path / "ref" / "state_dict.npz"), np.load(tmp_path / "dropin" / "state_dict.npz")
    assert sorted(sa.files) == sorted(sb.files)                       # checkpoint key names incl. `duffision_list`
    for k in sa.files:
        assert sa[k].shape == sb[k].shape, k
        # two Adam steps: every weight moved by ≈ lr·sign(g); elements whose gradient is rounding noise may differ by 2·lr
        upd_a, upd_b = sa[k].astype(np.float64), sb[k].astype(np.float64)
        assert np.abs(upd_a - upd_b).max() <= 2.5e-3, k
        assert np.linalg.norm(upd_a - upd_b) <= 0.02 * max(np.linalg.norm(upd_a), 1e-12), k


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "data", "uci")), reason="needs the reference checkout with its UCI data")
def test_reference_trainer_with_swapped_loss(tmp_path, lib):
    """the reference's trainer also drives ctgcn_b200.loss.NegativeSamplingLoss (constructed by train.get_loss from
    the loader's lil rows / json lists, called as loss_model([embeddings, batch_indices]), .backward(), .item()).  Its draws come
    from another generator than the reference's, so the runs are compared loosely: same files, same checkpoint keys, finite
    embeddings that stay close to the reference run's after two epochs."""
    pd = pytest.importorskip("pandas")
    _run("ref", tmp_path / "ref", "CGCN-C", 1, 2)
    _run("dropin_loss", tmp_path / "dropin", "CGCN-C", 1, 2)
    files = sorted(f for f in os.listdir(tmp_path / "ref") if f.endswith(".csv"))
    assert len(files) == 1 and files == sorted(f for f in os.listdir(tmp_path / "dropin") if f.endswith(".csv"))
    a = pd.read_csv(tmp_path / "ref" / files[0], sep="\t", index_col=0)
    b = pd.read_csv(tmp_path / "dropin" / files[0], sep="\t", index_col=0)
    assert a.shape == b.shape == (1899, 128) and np.isfinite(b.values).all()
    assert np.linalg.norm(a.values - b.values) / np.linalg.norm(a.values) < 0.4        # 0.195 with these seeds
    sa, sb = np.load(tmp_path / "ref" / "state_dict.npz"), np.load(tmp_path / "dropin" / "state_dict.npz")
    assert sorted(sa.files) == sorted(sb.files)
