"""Sparse on-disk formats: files written by the REFERENCE's own saver (committed
fixtures, tests/golden/make_sparse_golden.py) are read by the GPU loader bit-exactly, and files written here are
entry-for-entry what the reference writes (and are read back by the reference's own reader when /root/reference is
present, i.e. in the build container)."""
import importlib.util
import io
import os
import zipfile

import numpy as np
import pytest
import torch

REF = "/root/reference/pruning/final_pruning_script/global_storing_as sparse.py"


def _dense(golden_dir):
    z = np.load(os.path.join(golden_dir, "sparse_ref_dense.npz"))
    return {k: z[k] for k in z.files}


def test_writer_produces_the_reference_layout(golden_dir, tmp_path):
    """Same entries, same storage decision per tensor (sparse above 70 % zeros when smaller, else compressed), same
    indices / values / shapes as the file the reference's save_whisper_optimized wrote for the same state dict."""
    from openai_whisper_compression_b200 import sparse_store
    dense = _dense(golden_dir)
    out = str(tmp_path / "ours.zip")
    sparse_store.save_whisper_optimized({k: torch.from_numpy(v) for k, v in dense.items()}, out)
    with zipfile.ZipFile(out) as a, zipfile.ZipFile(os.path.join(golden_dir, "sparse_ref_optimized.zip")) as b:
        assert sorted(a.namelist()) == sorted(b.namelist())
        assert a.read("metadata.txt") == b.read("metadata.txt")
        for n in b.namelist():
            if n.endswith(".txt"):
                assert a.read(n) == b.read(n), n
            elif n.endswith(".npy"):
                np.testing.assert_array_equal(np.load(io.BytesIO(a.read(n))), np.load(io.BytesIO(b.read(n))))
            elif n.endswith(".npz"):
                np.testing.assert_array_equal(np.load(io.BytesIO(a.read(n)))["data"], np.load(io.BytesIO(b.read(n)))["data"])
        assert a.read("model.decoder.layers.0.fc1.weight/format.txt") == b"sparse"
        assert a.read("model.decoder.layers.0.fc2.weight/format.txt") == b"compressed"
        assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in a.infolist())


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree exists in the build container only")
def test_reference_reader_reads_our_file(golden_dir, tmp_path):
    from openai_whisper_compression_b200 import sparse_store
    spec = importlib.util.spec_from_file_location("ref_sparse", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    dense = _dense(golden_dir)
    out = str(tmp_path / "ours.zip")
    sparse_store.save_whisper_optimized({k: torch.from_numpy(v) for k, v in dense.items()}, out)
    back = ref.load_whisper_optimized(out)            # the reference's own loader, unchanged
    assert set(back) == set(dense)
    for k, v in dense.items():
        np.testing.assert_array_equal(back[k].numpy(), v)


def test_loader_has_no_cpu_path(golden_dir):
    from openai_whisper_compression_b200 import sparse_store
    with pytest.raises(RuntimeError):
        sparse_store.load_whisper_optimized(os.path.join(golden_dir, "sparse_ref_optimized.zip"), device="cpu")


@pytest.mark.gpu
def test_gpu_loader_reads_reference_files_bit_exact(golden_dir):
    from openai_whisper_compression_b200 import sparse_store
    dense = _dense(golden_dir)
    got = sparse_store.load_whisper_optimized(os.path.join(golden_dir, "sparse_ref_optimized.zip"), "cuda")
    coo = sparse_store.load_sparse_state_dict(os.path.join(golden_dir, "sparse_ref_coo.pt"), "cuda")
    assert set(got) == set(dense) == set(coo)
    for k, v in dense.items():
        assert got[k].is_cuda and got[k].dtype == torch.float32 and tuple(got[k].shape) == v.shape
        np.testing.assert_array_equal(got[k].cpu().numpy(), v)
        np.testing.assert_array_equal(coo[k].cpu().numpy(), v)


@pytest.mark.gpu
def test_gpu_loader_full_size_round_trip_and_quantize_from_device(tmp_path):
    """A 90 %-pruned d = 1280 / ffn = 5120 weight pair through the zip format: the GPU scatter reproduces the tensor
    exactly (sparsity-equality check of the reference, global_storing_as sparse.py:644-672), a corrupt index is
    refused, and the drop-in modules quantize straight from the device-resident tensors (zeros stay zeros)."""
    from openai_whisper_compression_b200 import bnb, sparse_store
    g = torch.Generator().manual_seed(2)
    lin = torch.nn.Linear(1280, 5120)
    with torch.no_grad():
        lin.weight[torch.rand(5120, 1280, generator=g) < 0.9] = 0
    path = str(tmp_path / "pruned.zip")
    sparse_store.save_whisper_optimized(lin, path)
    assert os.path.getsize(path) < 0.25 * lin.weight.numel() * 4
    sd = sparse_store.load_whisper_optimized(path, "cuda")
    assert torch.equal(sd["weight"].cpu(), lin.weight.detach()) and torch.equal(sd["bias"].cpu(), lin.bias.detach())
    assert int((sd["weight"] == 0).sum()) == int((lin.weight == 0).sum())
    m = bnb.Linear4bit(1280, 5120, bias=True, compute_dtype=torch.float16, compress_statistics=False, quant_type="nf4")
    m, _ = sparse_store.load_into(m, path, "cuda")
    wd = bnb.dequantize_4bit(m.weight.data, m.weight.quant_state)
    assert torch.all(wd[(lin.weight.detach() == 0).cuda()] == 0)
    with pytest.raises(ValueError):
        sparse_store._scatter(torch.tensor([5, 99], dtype=torch.int64), None, 0, torch.ones(2), 10, "cuda")
