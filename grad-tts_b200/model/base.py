"""BaseModule: same helpers as the reference (model/base.py:13-37): `.nparams`, `.relocate_input`."""
import torch


class BaseModule(torch.nn.Module):
    @property
    def nparams(self):
        """Number of trainable parameters (reference model/base.py:17-26)."""
        return sum(int(p.numel()) for p in self.parameters() if p.requires_grad)

    def relocate_input(self, x: list):
        """Move tensors to the module's device (reference model/base.py:29-37)."""
        device = next(self.parameters()).device
        for i, v in enumerate(x):
            if isinstance(v, torch.Tensor) and v.device != device:
                x[i] = v.to(device)
        return x
