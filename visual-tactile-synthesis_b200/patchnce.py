"""PatchNCELoss with the reference's class API (models/patchnce.py:6-55), forward and the explicit
gradient w.r.t. feat_q on one warp-shuffle CUDA kernel (feat_k is detached, as in the reference)."""
import torch
from torch import nn

from . import ops


class PatchNCELoss(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt

    def _batch(self):
        # patchnce.py:31-35: negatives come from the same image unless the flag pools the minibatch
        return 1 if self.opt.nce_includes_all_negatives_from_minibatch else self.opt.batch_size

    def forward(self, feat_q, feat_k):
        """-> per-patch loss [B*P] (reduction='none' cross entropy against the positive at index 0)."""
        if not feat_q.is_cuda:
            raise RuntimeError("PatchNCELoss (B200 path) needs CUDA tensors; there is no CPU fallback")
        loss, _ = ops.patchnce(feat_q.contiguous().float(), feat_k.detach().contiguous().float(), self._batch(), self.opt.nce_T)
        return loss

    def forward_backward(self, feat_q, feat_k, gscale):
        """loss and d(sum(loss) * gscale)/d feat_q in one launch (explicit backward)."""
        return ops.patchnce(feat_q.contiguous().float(), feat_k.detach().contiguous().float(), self._batch(), self.opt.nce_T,
                            want_grad=True, gscale=gscale)
