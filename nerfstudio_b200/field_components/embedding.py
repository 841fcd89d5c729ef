"""Per-image appearance embedding (mirror of nerfstudio/field_components/embedding.py:25-54).
A plain table lookup; state_dict key `embedding.weight` as in the reference."""
import torch
from torch import nn


class Embedding(nn.Module):
    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.embedding = torch.nn.Embedding(in_dim, out_dim)

    def get_out_dim(self) -> int:
        return self.out_dim

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)

    def forward(self, in_tensor):
        return self.embedding(in_tensor)
