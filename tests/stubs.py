"""Stand-ins for the out-of-scope modules either side of the hot path (encoder forward, StyleGAN3 generator)."""
import torch
import torch.nn as nn


class StubEncoder(nn.Module):
    """Returns seeded W+ `means` of the right shape instead of running the VGG encoder."""
    w_dim = 512
    num_ws = 16

    def __init__(self, sigma=0.14, seed=0):
        super().__init__()
        self.sigma, self.seed = sigma, seed
        self.anchor = nn.Parameter(torch.zeros(1))

    def means_for(self, B):
        g = torch.Generator().manual_seed(self.seed)
        return torch.randn(B, self.num_ws, self.w_dim, generator=g) * self.sigma

    def forward(self, x):
        m = self.means_for(x.shape[0]).to(self.anchor.device)
        return m + 0.01, m, torch.zeros_like(m)


class _Synthesis(nn.Module):
    def forward(self, w, noise_mode="const"):
        return w.mean(dim=1).reshape(w.shape[0], 1, 16, 32).expand(-1, 3, -1, -1).contiguous()


class StubGenerator(nn.Module):
    def __init__(self):
        super().__init__()
        self.synthesis = _Synthesis()
        self.scale = nn.Parameter(torch.ones(1))
