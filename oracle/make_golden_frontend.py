"""Golden vectors for the phoneme tensoriser (TEST INFRASTRUCTURE; runs only where /root/reference exists).

Calls the reference's own `ArticulatoryCombinedTextFrontend.string_to_tensor(..., input_phonemes=True)`
(Preprocessing/TextFrontend.py:213-288) on seeded phoneme strings -- base phonemes, stress marks, every modifier
character, unknown characters -- and stores the strings, the outputs, and the reference's lookup tables
(`generate_feature_table()`, `get_feature_to_index_lookup()`) in tests/golden/frontend.pt.

    python -m oracle.make_golden_frontend
"""
import io
import os
import random
from contextlib import redirect_stdout

import torch

from . import restate, shim


def make_strings(keys, n, seed):
    rng = random.Random(seed)
    mods = list(restate._PREVIOUS_PHONE_MODIFIERS)
    out = []
    for i in range(n):
        length = rng.randint(1, 60)
        chars = [rng.choice(keys)]                                   # a phoneme first: modifiers need a predecessor
        for _ in range(length):
            r = rng.random()
            if r < 0.70:
                chars.append(rng.choice(keys))
            elif r < 0.80:
                chars.append("ˈ")
            elif r < 0.95:
                chars.append(rng.choice(mods))
            elif r < 0.98:
                chars.append(rng.choice("ɚᵻ"))              # the two characters string_to_tensor rewrites
            else:
                chars.append(rng.choice("XQ7中\U0001F600"))       # unknown phonemes (skipped, printed by the reference)
        out.append("".join(chars))
    out += ["~", "aˈ", "ˈˈaːː", "aˈXːb", "hello woɹld~#"]
    return out


def main():
    shim.install(shim.FRONTEND_STUBS)
    from Preprocessing import TextFrontend
    from Preprocessing.articulatory_features import generate_feature_table, get_feature_to_index_lookup

    class Stub:
        pass
    stub = Stub()
    stub.phone_to_vector = generate_feature_table()
    f2i = dict(get_feature_to_index_lookup())
    keys = list(stub.phone_to_vector.keys())
    strings = make_strings(keys, 200, seed=11)
    outs = []
    with redirect_stdout(io.StringIO()):
        for s in strings:
            outs.append(TextFrontend.ArticulatoryCombinedTextFrontend.string_to_tensor(stub, s, input_phonemes=True))
    for s, ref in zip(strings, outs):                                # pin the restatement while we are here
        got = restate.string_to_tensor(s, stub.phone_to_vector, f2i)
        assert got.shape == ref.shape and torch.equal(got, ref), s
        assert bool(((ref == 0) | (ref == 1)).all())
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frontend.pt")
    torch.save({"phone_to_vector": {k: [float(x) for x in v] for k, v in stub.phone_to_vector.items()},
                "feature_to_index": f2i, "strings": strings,
                "outputs": [o.to(torch.uint8) for o in outs]}, path)                 # every value is 0 or 1
    print(f"wrote {path}: {len(strings)} strings, table {len(keys)} x {len(f2i)}")


if __name__ == "__main__":
    main()
