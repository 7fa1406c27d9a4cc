import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor for a single generator: draw on the generator's device."""
    gen_dev = generator.device if generator is not None else device
    x = torch.randn(shape, generator=generator, device=gen_dev, dtype=dtype)
    return x.to(device) if device is not None else x
