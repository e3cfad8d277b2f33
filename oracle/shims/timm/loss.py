"""Closed forms of timm 0.9.12 `SoftTargetCrossEntropy` and
`LabelSmoothingCrossEntropy` (published algorithm; call sites
/root/reference/model/loss.py:2,247,249).  Test infrastructure only."""
import torch
import torch.nn.functional as F


class SoftTargetCrossEntropy(torch.nn.Module):
    def forward(self, x, target):
        return torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1).mean()


class LabelSmoothingCrossEntropy(torch.nn.Module):
    def __init__(self, smoothing=0.1):
        super().__init__()
        assert smoothing < 1.0
        self.smoothing = smoothing
        self.confidence = 1.0 - smoothing

    def forward(self, x, target):
        logp = F.log_softmax(x, dim=-1)
        nll = -logp.gather(dim=-1, index=target.unsqueeze(1)).squeeze(1)
        smooth = -logp.mean(dim=-1)
        return (self.confidence * nll + self.smoothing * smooth).mean()
