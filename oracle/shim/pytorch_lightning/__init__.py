"""Stub: lvdm/models/ddpm3d.py only needs the LightningModule base class to be importable."""
import torch.nn as nn

LightningModule = nn.Module
