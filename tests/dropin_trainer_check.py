"""TEST INFRASTRUCTURE, launched by tests/test_dropin_trainer.py in a subprocess (it imports the reference's top-level modules
`train`, `models`, `layers`, … which must not leak into the pytest process).

    python tests/dropin_trainer_check.py {ref|dropin|dropin_loss} OUTDIR METHOD T EPOCHS

Runs the UNMODIFIED reference pipeline from /root/reference on the first T UCI snapshots — preprocessing, then
train.gnn_embedding(METHOD) with the shipped hyper-parameters — and copies the exported embeddings and the saved checkpoint to
OUTDIR.  In `dropin` mode ctgcn_b200.install_as_reference_modules() swaps the hot-path classes first, so the
reference's trainer drives ctgcn_b200's modules: constructors, state_dict, forward conventions, autograd, Adam.  There is no GPU
in the build container, so the CUDA entry points are replaced by the oracle-backed stand-in of tests/fake_backend.py; the
kernels themselves are checked against the same reference on the GPU (tests/test_parity_gpu.py, tests/test_train_gpu.py)."""
import json
import os
import random
import shutil
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ref_compat  # noqa: E402

ref_compat.apply()


# The reference's loss re-seeds Python's RNG from OS entropy on every call (`random.seed()`, metrics.py:72), which makes two runs
# of the reference itself differ.  For a run-to-run comparison argument-less re-seeding is ignored — in BOTH modes.
_seed = random.seed
random.seed = lambda *a, **k: _seed(*a, **k) if (a or k) else None


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def main():
    mode, outdir, method, T, epochs = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    torch.set_num_threads(2)
    cfg = json.load(open(os.path.join(REF, "config/uci.json")))
    tmp = tempfile.mkdtemp(prefix="dropin_")
    try:
        base = os.path.join(tmp, "uci")
        os.makedirs(os.path.join(base, "1.format"))
        os.makedirs(os.path.join(base, "nodes_set"))
        for f in sorted(os.listdir(os.path.join(REF, "data/uci/1.format")))[:T]:
            shutil.copy(os.path.join(REF, "data/uci/1.format", f), os.path.join(base, "1.format", f))
        shutil.copy(os.path.join(REF, "data/uci/nodes_set/nodes.csv"), os.path.join(base, "nodes_set/nodes.csv"))
        random.seed(0)
        np.random.seed(0)
        torch.manual_seed(0)
        from preprocessing import preprocess
        # every k-core method's embedding config reads CTGCN/ctgcn_cores and CTGCN/ctgcn_walk_pairs, which only the CTGCN-C
        # preprocessing entry writes (config/uci.json)
        preprocess("CTGCN-C", dict(cfg["preprocessing"]["CTGCN-C"], base_path=base, worker=-1))

        import models as ref_models
        if mode == "dropin_loss":      # the unsupervised loss too: rebound BEFORE train.py binds the name (train.py:8)
            import metrics as ref_metrics
            import fake_backend
            pkg = fake_backend.install(_Patch())
            pkg.install_as_reference_modules(ref_models, ref_metrics)
            assert ref_metrics.NegativeSamplingLoss is pkg.loss.NegativeSamplingLoss
        import train
        if mode == "dropin_loss":
            assert train.NegativeSamplingLoss is pkg.loss.NegativeSamplingLoss
        if mode == "dropin":
            import fake_backend
            pkg = fake_backend.install(_Patch())
            pkg.install_as_reference_modules(ref_models)
        if mode != "ref":
            assert ref_models.CTGCN is pkg.CTGCN and ref_models.CGCN is pkg.CGCN and sys.modules["layers"] is pkg.layers
        random.seed(1)
        np.random.seed(1)
        torch.manual_seed(1)
        args = dict(cfg["embedding"][method], base_path=base, duration=T, start_idx=0, end_idx=-1, epoch=epochs, use_cuda=False,
                    has_cuda=False, thread_num=2)
        train.gnn_embedding(method, args)
        os.makedirs(outdir, exist_ok=True)
        emb_dir = os.path.join(base, args["embed_folder"])
        for f in sorted(os.listdir(emb_dir)):
            shutil.copy(os.path.join(emb_dir, f), os.path.join(outdir, f))
        sd = torch.load(os.path.join(base, args["model_folder"], args["model_file"]))
        np.savez(os.path.join(outdir, "state_dict.npz"), **{k: v.numpy() for k, v in sd.items()})
        print("dropin_trainer_check OK", mode, method, sorted(os.listdir(outdir)))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
