import torch


class ResnetBlock2D(torch.nn.Module):
    """isinstance() target only."""
