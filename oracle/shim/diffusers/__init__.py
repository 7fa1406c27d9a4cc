"""Minimal stand-in for the parts of `diffusers` that the reference's scheduler/pipeline import
(scheduler/t2v_turbo_scheduler.py:1-30, pipeline/t2v_turbo_vc2_pipeline.py:1-10).  diffusers is not
installed in this image; the VC2 path uses it only for base classes and `randn_tensor` (no arithmetic
beyond torch.randn).  TEST INFRASTRUCTURE ONLY: used by oracle/make_goldens.py in the authoring
container to run the UNMODIFIED reference; never imported by the product."""
import contextlib

import torch
import torch.nn as nn

from . import logging  # noqa: F401
from .configuration_utils import ConfigMixin  # noqa: F401


class SchedulerMixin:
    pass


class _Bar:
    def update(self, *a, **k):
        pass


class DiffusionPipeline(nn.Module):
    def register_modules(self, **kw):
        for k, v in kw.items():
            if isinstance(v, nn.Module):
                setattr(self, k, v)
            else:
                object.__setattr__(self, k, v)

    @property
    def _execution_device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    @property
    def dtype(self):
        for p in self.parameters():
            return p.dtype
        return torch.float32

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        yield _Bar()
