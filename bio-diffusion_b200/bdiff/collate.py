"""Packed training collation on the device — replaces the padded PyG batches of the reference's QM9 pipeline
(`ProcessedDataset._featurize_as_graph`, datamodules/components/edm_dataset.py:187-216, + the PyG collater) and
`prepare_context` (datamodules/components/edm/utils.py:333-382).

The padded dataset tensors are uploaded once (`PackedDataset`); `collate(idx)` gathers the present atoms of the selected
molecules into the packed layout the denoiser consumes — the reference batch restricted to `mask == True`, bit-exact — in
two kernels and one 4-byte device-to-host read (the row count).  The returned `PackedBatch` carries the attributes the
reference's model code reads from a PyG batch (`batch`, `mask`, `x`, `one_hot`, `charges`, `props_context`,
`num_graphs`, `num_nodes_present`).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib


class PackedBatch:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __getitem__(self, key):
        return getattr(self, key)

    @property
    def num_nodes(self):
        return int(self.batch.shape[0])


class PackedDataset:
    """positions [M,P,3], charges [M,P] (0 = no atom), one_hot [M,P,A] and any per-molecule float properties [M]."""

    def __init__(self, data: Dict[str, torch.Tensor], device: torch.device, properties: Sequence[str] = ()):
        if torch.device(device).type != "cuda":
            raise _lib.BdiffError("PackedDataset lives on a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.positions = data["positions"].to(self.device, torch.float32).contiguous()
        self.charges = data["charges"].to(self.device, torch.int32).contiguous()
        self.one_hot = data["one_hot"].to(self.device, torch.uint8).contiguous()
        self.m, self.pad = self.charges.shape
        self.num_types = int(self.one_hot.shape[-1])
        self.prop_names: List[str] = list(properties)
        self.props = (torch.stack([data[k].to(self.device, torch.float32).reshape(self.m) for k in self.prop_names])
                      .contiguous() if self.prop_names else None)

    def collate(self, idx: torch.Tensor, conditioning: Sequence[str] = (),
                property_norms: Optional[Dict[str, Dict[str, torch.Tensor]]] = None) -> PackedBatch:
        lib = _lib.load()
        dev = self.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        idx = idx.to(dev, torch.int64).contiguous()
        b = int(idx.shape[0])
        counts = torch.empty(b, dtype=torch.int32, device=dev)
        rc = lib.bdiff_collate_count(st, C.c_void_p(self.charges.data_ptr()), C.c_void_p(idx.data_ptr()), b, self.pad,
                                     C.c_void_p(counts.data_ptr()))
        if rc != 0:
            raise _lib.BdiffError(f"bdiff_collate_count failed with code {rc}")
        off = torch.zeros(b + 1, dtype=torch.int32, device=dev)
        off[1:] = torch.cumsum(counts, 0)
        n = int(off[-1].item())                                   # the one host read of a collation
        x = torch.empty((n, 3), device=dev)
        oh = torch.empty((n, self.num_types), device=dev)
        ch = torch.empty((n, 1), device=dev)
        bi = torch.empty(n, dtype=torch.int64, device=dev)
        rc = lib.bdiff_collate_packed(st, C.c_void_p(self.positions.data_ptr()), C.c_void_p(self.charges.data_ptr()),
                                      C.c_void_p(self.one_hot.data_ptr()), C.c_void_p(idx.data_ptr()),
                                      C.c_void_p(off.data_ptr()), b, self.pad, self.num_types, C.c_void_p(x.data_ptr()),
                                      C.c_void_p(oh.data_ptr()), C.c_void_p(ch.data_ptr()), C.c_void_p(bi.data_ptr()))
        if rc != 0:
            raise _lib.BdiffError(f"bdiff_collate_packed failed with code {rc}")
        ctx = None
        if conditioning:
            sel = [self.prop_names.index(k) for k in conditioning]
            props = self.props[sel].contiguous()
            mean = torch.stack([property_norms[k]["mean"].reshape(()) for k in conditioning]).to(dev, torch.float32)
            mad = torch.stack([property_norms[k]["mad"].reshape(()) for k in conditioning]).to(dev, torch.float32)
            ctx = torch.empty((n, len(sel)), device=dev)
            rc = lib.bdiff_prepare_context(st, C.c_void_p(props.data_ptr()), C.c_void_p(idx.data_ptr()),
                                           C.c_void_p(bi.data_ptr()), C.c_void_p(mean.data_ptr()), C.c_void_p(mad.data_ptr()),
                                           self.m, n, len(sel), C.c_void_p(ctx.data_ptr()))
            if rc != 0:
                raise _lib.BdiffError(f"bdiff_prepare_context failed with code {rc}")
        return PackedBatch(batch=bi, mask=torch.ones(n, dtype=torch.bool, device=dev), x=x, one_hot=oh, charges=ch,
                           index=idx.unsqueeze(-1), props_context=ctx, num_graphs=b, num_nodes_present=counts.to(torch.int64))
