"""On-disk / wire formats around the hot path, so that real checkpoints and preprocessed datasets drop in.

* `unet_lora.pt` — flat `[up_0, down_0, ...]` list: `lora.merge_lora` (inference) / `lora_train.LoraArena.load_list`.
* VideoCrafter2 `model.ckpt` key space: `pipeline.LatentVideoModel.load_vc2_checkpoint`.
* `unet.pt` / `unet_mg.pt` — the T2V-Turbo-v2 full-UNet state dicts (`predict.py:47-56`, `app.py`): a plain `state_dict()` of
  `UNetModel(time_cond_proj_dim=256[, motion_cond_proj_dim=256])`, loaded strictly.
* the v2 preprocessed-latent sample (`preprocess_scripts/preprocess_with_motion_prior.py:392-405`, read back by
  `train_latent_t2v_turbo_v2.py:978-985`): a pickled dict of fp16 CPU tensors
  `{index, z_t, cond_teacher_out, uncond_teacher_out, score, z_example, z_example_prev, prompt_emb}`.
"""
from __future__ import annotations

import io
import pickle

import torch

V2_SAMPLE_KEYS = ("index", "z_t", "cond_teacher_out", "uncond_teacher_out", "score", "z_example", "z_example_prev", "prompt_emb")
_V2_LATENT_KEYS = ("z_t", "cond_teacher_out", "uncond_teacher_out", "score", "z_example", "z_example_prev")


def load_unet_weights(unet, ckpt, strict: bool = True):
    """`unet.pt` / `unet_mg.pt` (path, file object or an already loaded state dict) into a B200 `UNetModel`.  The v2
    checkpoints are full fine-tuned UNets, so the load is strict: a key mismatch means the UNet was built without
    `time_cond_proj_dim` / `motion_cond_proj_dim` (predict.py:49-50) and raises with the offending keys."""
    sd = ckpt if isinstance(ckpt, dict) else torch.load(ckpt, map_location="cpu", weights_only=True)
    sd = sd.get("state_dict", sd)
    own = unet.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    if strict and (missing or unexpected):
        raise RuntimeError(f"load_unet_weights: missing {missing[:6]}{' ...' if len(missing) > 6 else ''}, unexpected "
                           f"{unexpected[:6]}{' ...' if len(unexpected) > 6 else ''} (build the UNet with time_cond_proj_dim=256 and, "
                           "for unet_mg.pt, motion_cond_proj_dim=256)")
    bad = [k for k in own if k in sd and tuple(sd[k].shape) != tuple(own[k].shape)]
    if bad:
        raise RuntimeError(f"load_unet_weights: shape mismatch for {bad[:6]}")
    unet.load_state_dict(sd, strict=strict)     # (drops the packed bf16 operands and bumps weight_generation)
    return missing, unexpected


def check_v2_sample(sample: dict, frames: int | None = None) -> dict:
    """Validate one v2 preprocessed sample against the writer's schema; returns it unchanged."""
    keys = set(sample)
    if keys != set(V2_SAMPLE_KEYS):
        raise ValueError(f"v2 sample: keys {sorted(keys ^ set(V2_SAMPLE_KEYS))} differ from the schema {V2_SAMPLE_KEYS}")
    ref = sample["z_t"]
    for k in _V2_LATENT_KEYS:
        t = sample[k]
        if not torch.is_tensor(t) or t.dtype != torch.float16 or t.device.type != "cpu":
            raise ValueError(f"v2 sample: {k} must be a CPU fp16 tensor")
        if t.dim() != 4 or t.shape != ref.shape:
            raise ValueError(f"v2 sample: {k} has shape {tuple(t.shape)}, z_t has {tuple(ref.shape)} ([C, T, h, w] expected)")
    if frames is not None and ref.shape[1] != frames:
        raise ValueError(f"v2 sample: {ref.shape[1]} frames, expected {frames}")
    pe = sample["prompt_emb"]
    if not torch.is_tensor(pe) or pe.dtype != torch.float16 or pe.dim() != 2:
        raise ValueError("v2 sample: prompt_emb must be an fp16 [tokens, dim] tensor")
    idx = sample["index"]
    if not (torch.is_tensor(idx) and idx.numel() == 1):
        raise ValueError("v2 sample: index must be a one-element tensor (the DDIM timestep index)")
    return sample


def dumps_v2_sample(index, z_t, cond_teacher_out, uncond_teacher_out, score, z_example, z_example_prev, prompt_emb) -> bytes:
    """Serialise one sample exactly as preprocess_with_motion_prior.py:392-403 does (fp16, CPU, pickle)."""
    to_save = {"index": index, "z_t": z_t, "cond_teacher_out": cond_teacher_out, "uncond_teacher_out": uncond_teacher_out,
               "score": score, "z_example": z_example, "z_example_prev": z_example_prev, "prompt_emb": prompt_emb}
    to_save = {k: (v.to(torch.float16) if k != "index" else torch.as_tensor(v)).detach().cpu() for k, v in to_save.items()}
    return pickle.dumps(check_v2_sample(to_save))


class _TensorOnlyUnpickler(pickle.Unpickler):
    """The samples come from object storage: only torch tensor reconstruction and plain containers are allowed."""
    _OK = {("torch._utils", "_rebuild_tensor_v2"), ("torch", "HalfStorage"), ("torch", "LongStorage"), ("torch", "FloatStorage"),
           ("torch.storage", "_load_from_bytes"), ("collections", "OrderedDict"), ("torch", "Size"), ("torch", "float16"),
           ("torch", "int64"), ("torch.serialization", "_get_layout"), ("torch", "device")}

    def find_class(self, module, name):
        if (module, name) in self._OK:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"v2 sample: refusing to unpickle {module}.{name}")


def loads_v2_sample(data: bytes, frames: int | None = None) -> dict:
    return check_v2_sample(_TensorOnlyUnpickler(io.BytesIO(data)).load(), frames)


def collate_v2_samples(samples, device=None, dtype=torch.bfloat16) -> dict:
    """Batch the samples the way train_latent_t2v_turbo_v2.py:978-985 consumes them: latents [B, C, T, h, w] and prompt
    embeddings [B, tokens, dim] in the training dtype on `device`, `index` as a long vector."""
    out = {k: torch.stack([s[k] for s in samples]).to(device=device, dtype=dtype) for k in _V2_LATENT_KEYS + ("prompt_emb",)}
    out["index"] = torch.stack([s["index"].reshape(()) for s in samples]).to(device=device, dtype=torch.long)
    return out
