"""Input contract of the hot path: NestedTensor + nested_tensor_from_tensor_list (reference:
util/misc.py:294-339).  Stand-alone, only what LWDETR.forward consumes is defined here; the reference's
distributed / logging helpers are out of scope.

When the reference tree is on sys.path behind this package, the reference's scripts expect the FULL
`util.misc` (`utils.init_distributed_mode`, `is_main_process`, ... demo/demo.py:171, main.py:186): in that case this
module executes the reference's own util/misc.py in its place, so callers see exactly the reference's definitions
(including its NestedTensor class) and nothing is re-implemented."""
import os as _os
import sys as _sys


def _reference_misc():
    here = _os.path.dirname(_os.path.abspath(__file__))
    for entry in _sys.path:
        cand = _os.path.join(entry or ".", "util", "misc.py")
        if _os.path.isfile(cand) and _os.path.dirname(_os.path.abspath(cand)) != here and \
                _os.path.isfile(_os.path.join(entry or ".", "models", "lwdetr.py")):
            return cand
    return None


_REF_MISC = _reference_misc()
if _REF_MISC is not None:
    with open(_REF_MISC) as _f:
        exec(compile(_f.read(), _REF_MISC, "exec"), globals())

if _REF_MISC is None:
    from typing import List, Optional

    import torch
    from torch import Tensor


    class NestedTensor(object):
        """A batch of images padded to a common size plus the padding mask (True = padded pixel)."""

        def __init__(self, tensors: Tensor, mask: Optional[Tensor]):
            self.tensors = tensors
            self.mask = mask

        def to(self, device):
            m = self.mask.to(device) if self.mask is not None else None
            return NestedTensor(self.tensors.to(device), m)

        def decompose(self):
            return self.tensors, self.mask

        def __repr__(self):
            return str(self.tensors)


    def nested_tensor_from_tensor_list(tensor_list):
        """List of [3,H,W] tensors (or an already batched [B,3,H,W] tensor) -> NestedTensor, zero padded to the
        largest H and W with mask=True on the padding."""
        if isinstance(tensor_list, Tensor):
            if tensor_list.dim() != 4:
                raise ValueError("expected a [B, 3, H, W] tensor")
            b, _, h, w = tensor_list.shape
            return NestedTensor(tensor_list, torch.zeros((b, h, w), dtype=torch.bool, device=tensor_list.device))
        if len(tensor_list) == 0 or tensor_list[0].dim() != 3:
            raise ValueError("not supported")
        hmax = max(int(t.shape[1]) for t in tensor_list)
        wmax = max(int(t.shape[2]) for t in tensor_list)
        first = tensor_list[0]
        batch = torch.zeros((len(tensor_list), first.shape[0], hmax, wmax), dtype=first.dtype, device=first.device)
        mask = torch.ones((len(tensor_list), hmax, wmax), dtype=torch.bool, device=first.device)
        for i, img in enumerate(tensor_list):
            batch[i, :, : img.shape[1], : img.shape[2]].copy_(img)
            mask[i, : img.shape[1], : img.shape[2]] = False
        return NestedTensor(batch, mask)
