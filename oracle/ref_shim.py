"""ORACLE tooling (build container only): import the UNMODIFIED reference from /root/reference.

`gill/models.py:3` imports `diffusers.StableDiffusionPipeline`, which is not installed here; a one-attribute stub
module lets the file import unchanged. `OPTForCausalLM.from_pretrained` / `CLIPVisionModel.from_pretrained` /
`utils.get_feature_extractor_for_model` need the HF hub, so they are patched to config-built, seeded, random-weight
models. Nothing here is used at run time on the GPU box (it has no /root/reference); it only produces the fixtures
under tests/golden/ (see oracle/make_golden.py).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("GILL_REFERENCE", "/root/reference")
CKPT_DIR = os.path.join(REF_ROOT, "checkpoints", "gill_opt")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "gill"))


def import_reference():
    """Returns the reference's (layers, models, utils) modules."""
    if "diffusers" not in sys.modules:
        stub = types.ModuleType("diffusers")
        stub.StableDiffusionPipeline = type("StableDiffusionPipeline", (), {})
        sys.modules["diffusers"] = stub
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from gill import layers, models, utils  # type: ignore

    return layers, models, utils


def load_real_state_dict():
    """checkpoints/gill_opt/pretrained_ckpt.pth.tar with the DDP prefix stripped (gill/models.py:880-884)."""
    import torch

    ck = torch.load(os.path.join(CKPT_DIR, "pretrained_ckpt.pth.tar"), map_location="cpu")
    return {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}


def mapper_state_dict(sd=None):
    sd = sd or load_real_state_dict()
    pre = "model.gen_text_hidden_fcs.0."
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


def reference_mapper(dtype=None):
    """The reference's own TextFcLayer (gill/models.py:109-116 construction) with the shipped weights."""
    import torch

    layers, _, _ = import_reference()
    m = layers.TextFcLayer(4096, 768, num_input_tokens=8, num_output_tokens=77, mode="gill_mapper")
    m.load_state_dict(mapper_state_dict(), strict=True)
    m.eval()
    return m.to(dtype or torch.float32)
