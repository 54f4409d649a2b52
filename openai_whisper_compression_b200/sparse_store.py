"""Sparse on-disk formats of the pruned models with a GPU-side loader.

Two formats, both the reference's own:

  * the "optimized" zip (pruning/final_pruning_script/global_storing_as sparse.py:287-407 writer, :410-485 reader):
    ZIP_DEFLATED level 9; ``metadata.txt``; per state-dict entry ``<name>/format.txt`` = ``sparse`` (more than 70 % zeros
    and indices + values smaller than the dense tensor: ``shape.txt``, ``dtype.txt``, ``indices.npy`` = flat int64
    positions of the non-zeros, ``values.npy``) or ``compressed`` (``data.npz`` from numpy.savez_compressed);
  * the COO state dict (pruning+quantization/bnb_implementation.py:386-486): ``torch.save`` of a dict whose entries
    with more than 30 % zeros were converted with ``.to_sparse()``.

Files are interchangeable with the reference's (same entry names and encodings; tests read a file its writer produced
and hand ours to its reader).  What changes is where the dense tensor is rebuilt: the reference does
``dense[indices] = values`` on the host and uploads 4 bytes per parameter; here indices and values go to the GPU as they
are (for a 90 %-sparse tensor 1.2 bytes per parameter over PCIe instead of 4) and ``wq_scatter_dense_f32`` scatters them
into the zero-filled tensor there, from where the drop-in modules quantize them in place (``load_into``).
There is no CPU reconstruction path: a non-CUDA device raises.
"""
from __future__ import annotations

import ast
import io
import zipfile
from collections import OrderedDict
from typing import Dict

import numpy as np
import torch

from . import _lib
from . import functional as F

SPARSE_MIN_SPARSITY = 70.0      # per cent; the reference's threshold for the (indices, values) form


def _npy(a: np.ndarray) -> bytes:
    b = io.BytesIO()
    np.save(b, a)
    return b.getvalue()


def save_whisper_optimized(model_or_state, output_path: str) -> float:
    """Write the state dict in the reference's zip layout; returns the file size in MB."""
    state = model_or_state.state_dict() if hasattr(model_or_state, "state_dict") else model_or_state
    with zipfile.ZipFile(output_path, "w", compression=zipfile.ZIP_DEFLATED, compresslevel=9) as zf:
        zf.writestr("metadata.txt", str({"model_type": "whisper", "format_version": "1.0", "compression": "zip_deflate"}))
        for name, t in state.items():
            a = t.detach().cpu().numpy()
            flat = a.reshape(-1)
            nz = np.flatnonzero(flat)
            sparsity = 100.0 * (flat.size - nz.size) / flat.size if flat.size else 0.0
            if sparsity > SPARSE_MIN_SPARSITY and nz.size * 8 < flat.size * 4:
                zf.writestr(f"{name}/format.txt", "sparse")
                zf.writestr(f"{name}/shape.txt", str(a.shape))
                zf.writestr(f"{name}/dtype.txt", str(a.dtype))
                zf.writestr(f"{name}/indices.npy", _npy(nz.astype(np.int64)))
                zf.writestr(f"{name}/values.npy", _npy(flat[nz]))
                continue
            b = io.BytesIO()
            np.savez_compressed(b, data=a)
            zf.writestr(f"{name}/format.txt", "compressed")
            zf.writestr(f"{name}/data.npz", b.getvalue())
    import os
    return os.path.getsize(output_path) / (1024 * 1024)


def _scatter(idx0: torch.Tensor, idx1, cols: int, vals: torch.Tensor, n_out: int, device) -> torch.Tensor:
    """Dense fp32 [n_out] on `device` from host index / value arrays (pinned staging, one scatter launch)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("sparse_store rebuilds tensors on a CUDA device (sm_100a library); there is no CPU path")
    out = torch.empty((n_out,), dtype=torch.float32, device=dev)
    err = torch.zeros((1,), dtype=torch.int32, device=dev)
    i0 = idx0.contiguous().pin_memory().to(dev, non_blocking=True)
    i1 = None if idx1 is None else idx1.contiguous().pin_memory().to(dev, non_blocking=True)
    v = vals.to(torch.float32).contiguous().pin_memory().to(dev, non_blocking=True)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().wq_scatter_dense_f32(i0.data_ptr(), None if i1 is None else i1.data_ptr(),
                                                    i0.element_size(), int(cols), v.data_ptr(), v.numel(),
                                                    out.data_ptr(), n_out, err.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream), "wq_scatter_dense_f32")
    F.STATS.launches += 1
    if int(err.item()):
        raise ValueError("sparse checkpoint holds an index outside its tensor")
    return out


def load_whisper_optimized(model_path: str, device="cuda") -> "OrderedDict[str, torch.Tensor]":
    """Read the reference's zip layout into a float32 state dict on `device` (the reference's loader returns float32
    for every entry, global_storing_as sparse.py:474,483)."""
    state: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    dev = torch.device(device)
    with zipfile.ZipFile(model_path, "r") as zf:
        names = zf.namelist()
        params = []
        for n in names:             # keep the file's order (the reference iterates a set)
            if "/" in n:
                p = n.rsplit("/", 1)[0]
                if p not in params:
                    params.append(p)
        for p in params:
            if f"{p}/format.txt" not in names:
                continue
            fmt = zf.read(f"{p}/format.txt").decode("utf-8").strip()
            if fmt == "sparse":
                shape = tuple(ast.literal_eval(zf.read(f"{p}/shape.txt").decode("utf-8").strip()))
                idx = np.load(io.BytesIO(zf.read(f"{p}/indices.npy")))
                val = np.load(io.BytesIO(zf.read(f"{p}/values.npy")))
                n_out = int(np.prod(shape)) if len(shape) else 1
                if idx.dtype not in (np.int32, np.int64):
                    idx = idx.astype(np.int64)
                dense = _scatter(torch.from_numpy(idx), None, 0, torch.from_numpy(np.ascontiguousarray(val)), n_out, dev)
                state[p] = dense.view(shape)
            elif fmt == "compressed":
                a = np.load(io.BytesIO(zf.read(f"{p}/data.npz")))["data"]
                if dev.type != "cuda":
                    raise RuntimeError("sparse_store loads onto a CUDA device; there is no CPU path")
                state[p] = torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32).to(dev)
            else:
                raise ValueError(f"unknown storage format {fmt!r} for {p}")
    return state


def load_sparse_state_dict(path_or_state, device="cuda") -> Dict[str, torch.Tensor]:
    """The COO state dict of the reference's save_sparse_model (bnb_implementation.py:386-486): sparse entries are
    densified on the GPU from their indices / values, dense ones are uploaded."""
    sd = torch.load(path_or_state, map_location="cpu", weights_only=False) if isinstance(path_or_state, (str, bytes, io.IOBase)) \
        else path_or_state
    out: Dict[str, torch.Tensor] = OrderedDict()
    dev = torch.device(device)
    for k, t in sd.items():
        if isinstance(t, torch.Tensor) and t.layout == torch.sparse_coo:
            t = t.coalesce()
            ind, val = t.indices(), t.values()
            shape = tuple(t.shape)
            n_out = int(np.prod(shape))
            if ind.shape[0] == 1:
                dense = _scatter(ind[0], None, 0, val, n_out, dev)
            elif ind.shape[0] == 2:
                dense = _scatter(ind[0], ind[1], shape[1], val, n_out, dev)
            else:       # flatten leading dimensions into a row index
                strides = np.cumprod((1,) + shape[:0:-1])[::-1]
                flat = sum(ind[i] * int(strides[i]) for i in range(ind.shape[0]))
                dense = _scatter(flat, None, 0, val, n_out, dev)
            out[k] = dense.view(shape).to(val.dtype)
        else:
            if dev.type != "cuda":
                raise RuntimeError("sparse_store loads onto a CUDA device; there is no CPU path")
            out[k] = t.to(dev) if isinstance(t, torch.Tensor) else t
    return out


def load_into(model: torch.nn.Module, path: str, device="cuda", strict: bool = True):
    """Rebuild the pruned weights on the GPU and load them into `model` (moved to `device` first): with the drop-in
    modules already swapped in, their load_state_dict / .to(device) hooks quantize from the device-resident dense
    tensors -- the dense fp32 model never exists on the host."""
    state = load_whisper_optimized(path, device) if zipfile.is_zipfile(path) else load_sparse_state_dict(path, device)
    missing = model.load_state_dict(state, strict=strict)
    return model.to(device), missing
