"""Dump the state_dict key/shape manifests of the reference's train-side models.

TEST INFRASTRUCTURE.  Run in the authoring container (needs the live reference):
    python -m oracle.make_manifest
Writes oracle/manifests/{toucantts,hifigan,bigvgan}.json.  The manifests are the
"same state_dict layout" contract: key -> shape, dtype and
(for parameters that share storage, e.g. the PostFlow WN layers shared within
groups of 4 coupling blocks, Glow.py:325-327) the canonical key they alias.
"""
import json
import os

import torch

from oracle import shim


def manifest_of(module):
    sd = module.state_dict()
    seen = {}
    out = {}
    for key, value in sd.items():
        entry = {"shape": list(value.shape), "dtype": str(value.dtype).replace("torch.", "")}
        ptr = (value.data_ptr(), tuple(value.shape)) if value.numel() > 0 else None
        if ptr is not None and ptr in seen:
            entry["alias"] = seen[ptr]
        elif ptr is not None:
            seen[ptr] = key
        out[key] = entry
    return out


def main():
    cls = shim.reference_classes()
    torch.manual_seed(0)
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "manifests")
    os.makedirs(here, exist_ok=True)
    for name, ctor in (("toucantts", cls["TrainToucanTTS"]), ("hifigan", cls["TrainHiFiGAN"]),
                       ("bigvgan", cls["TrainBigVGAN"])):
        man = manifest_of(ctor())
        with open(os.path.join(here, name + ".json"), "w") as f:
            json.dump(man, f, indent=0, separators=(",", ":"))
        print(name, len(man), "keys,", sum(1 for v in man.values() if "alias" in v), "aliases")


if __name__ == "__main__":
    main()
