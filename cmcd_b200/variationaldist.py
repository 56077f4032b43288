"""Diagonal-Gaussian variational distribution q -- parameter container only.

Mirrors /root/reference/src/variationaldist.py:4-13 and src/vardist/diag_gauss.py:20-23
(``initialize``).  Sampling (``sample_rep``), ``log_prob`` and the q-score are evaluated
inside the fused CUDA bridge kernel (cmcd_b200/csrc/bridge_fwd.cu), not here.
"""
import math

import torch


def initialize(dim, init_sigma=1.0, device=None, dtype=torch.float32):
    return {"mean": torch.zeros(dim, dtype=dtype, device=device),
            "logdiag": torch.ones(dim, dtype=dtype, device=device) * math.log(init_sigma)}
