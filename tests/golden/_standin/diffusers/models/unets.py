import torch


class UNet2DModel(torch.nn.Module):
    """isinstance() target only."""
