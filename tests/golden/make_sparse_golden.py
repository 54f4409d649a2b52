"""Golden fixture for the sparse on-disk formats, written by the REFERENCE's own code:
imports /root/reference/pruning/final_pruning_script/"global_storing_as sparse.py" (save_whisper_optimized) and
pruning+quantization/bnb_implementation.py's COO convention (param.to_sparse() for > 30 % zeros, torch.save) and stores a
small pruned state dict in both formats next to the dense arrays:

    python tests/golden/make_sparse_golden.py        # needs /root/reference (this container only)

Outputs: tests/golden/sparse_ref_optimized.zip, tests/golden/sparse_ref_coo.pt, tests/golden/sparse_ref_dense.npz
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pruning/final_pruning_script/global_storing_as sparse.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_sparse", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.RandomState(11)

    def pruned(shape, sparsity):
        w = rng.randn(*shape).astype(np.float32) * 0.02
        w[rng.rand(*shape) < sparsity] = 0
        return torch.from_numpy(w)

    state = {"model.decoder.layers.0.fc1.weight": pruned((96, 64), 0.9),        # sparse form
             "model.decoder.layers.0.fc1.bias": pruned((96,), 0.0),             # compressed form
             "model.decoder.layers.0.fc2.weight": pruned((64, 96), 0.5),        # not sparse enough: compressed
             "model.encoder.conv1.weight": pruned((8, 4, 3), 0.8),              # 3-d, sparse form
             "proj_out.weight": pruned((130, 64), 0.95)}

    class Holder:
        def state_dict(self):
            return state

    ref.save_whisper_optimized(Holder(), os.path.join(HERE, "sparse_ref_optimized.zip"))
    coo = {k: (v.to_sparse() if v.dim() > 0 and torch.sum(v == 0) > 0.3 * v.numel() else v) for k, v in state.items()}
    torch.save(coo, os.path.join(HERE, "sparse_ref_coo.pt"))
    np.savez_compressed(os.path.join(HERE, "sparse_ref_dense.npz"), **{k: v.numpy() for k, v in state.items()})
    # and back through the reference's reader as a self-check of the fixture
    back = ref.load_whisper_optimized(os.path.join(HERE, "sparse_ref_optimized.zip"))
    for k, v in state.items():
        assert torch.equal(back[k], v), k
    print("wrote sparse golden fixtures:", {k: tuple(v.shape) for k, v in state.items()})


if __name__ == "__main__":
    sys.exit(main())
