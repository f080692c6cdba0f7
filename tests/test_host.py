"""CPU: host-side logic -- state_dict layouts, the C-ABI library's exported symbols, loud failure
without a GPU.  No compute calls here."""
import ctypes
import json
import math
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layouts_equal_reference_manifests():
    from ims_toucan_prosody_variance_b200 import layouts
    for name, fn in (("hifigan", layouts.hifigan_layout), ("bigvgan", layouts.bigvgan_layout),
                     ("toucantts", layouts.toucantts_layout)):
        with open(os.path.join(ROOT, "oracle", "manifests", name + ".json")) as f:
            man = json.load(f)
        lay, alias = fn()
        assert list(lay) == list(man)
        for k, meta in man.items():
            assert tuple(meta["shape"]) == tuple(lay[k]), k
            assert meta.get("alias") == alias.get(k), k


def test_library_exports_every_declared_symbol():
    from ims_toucan_prosody_variance_b200 import _lib
    header = open(os.path.join(ROOT, "include", "toucan_b200.h")).read()
    declared = set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", header))
    declared -= {"tb200_conv1d_params", "tb200_respair_params"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/toucan_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert _lib.load().tb200_version() == int(re.search(r"#define TB200_VERSION (\d+)", header).group(1))


@pytest.mark.parametrize("struct,cls", [("tb200_conv1d_params", "Conv1dParams"), ("tb200_respair_params", "RespairParams")])
def test_params_struct_matches_header(struct, cls):
    """Field names, order and C types of the ctypes mirrors equal the structs of include/toucan_b200.h."""
    from ims_toucan_prosody_variance_b200 import _lib
    header = open(os.path.join(ROOT, "include", "toucan_b200.h")).read()
    body = header[header.index("typedef struct %s {" % struct):header.index("} %s;" % struct)]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split("{", 1)[1].split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        m = re.match(r"^(const\s+)?(void|float|int32_t|int64_t)\s*(.*)$", stmt)
        base = m.group(2)
        for n in m.group(3).split(","):
            n = n.strip()
            ptr = n.startswith("*")
            fields.append((n.lstrip("*").strip(), "ptr" if ptr else base))
    ctype = {"ptr": ctypes.c_void_p, "float": ctypes.c_float, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64}
    mirror = getattr(_lib, cls)._fields_
    assert [f[0] for f in mirror] == [f[0] for f in fields]
    assert [f[1] for f in mirror] == [ctype[f[1]] for f in fields]


def test_module_state_dict_roundtrip(tmp_path):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    for kind, cls in (("hifigan", tb.HiFiGANGenerator), ("bigvgan", tb.BigVGAN)):
        sd = factory.make_state_dict(kind, 7)
        path = os.path.join(tmp_path, kind + ".pt")
        torch.save({"generator": sd}, path)
        m = cls(path)
        got = m.state_dict()
        assert list(got) == list(sd)
        assert all(torch.equal(got[k], sd[k]) for k in sd)


def test_bigvgan_accepts_checkpoint_without_filter_buffers(tmp_path):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    sd = {k: v for k, v in factory.make_state_dict("bigvgan", 7).items() if not k.endswith(".filter")}
    path = os.path.join(tmp_path, "nofilt.pt")
    torch.save({"generator": sd}, path)
    assert list(tb.BigVGAN(path).state_dict()) == list(sd)


def test_no_cpu_fallback(tmp_path):
    """On a machine without CUDA the product path must fail loudly, not compute on the CPU."""
    import ims_toucan_prosody_variance_b200 as tb
    from ims_toucan_prosody_variance_b200._lib import EngineError
    from oracle import factory
    sd = factory.make_state_dict("hifigan", 7)
    path = os.path.join(tmp_path, "h.pt")
    torch.save({"generator": sd}, path)
    m = tb.HiFiGANGenerator(path)
    with pytest.raises(EngineError):
        m(torch.zeros(80, 10))
    with pytest.raises(EngineError):
        tb.ops.ConvLayer(torch.zeros(4, 4, 1))


def test_fold_weight_norm_matches_torch():
    from ims_toucan_prosody_variance_b200 import layouts
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(6, 5, 3))
    with torch.no_grad():
        conv.weight_g.mul_(1.7)
    sd = {"c." + k: v.detach().clone() for k, v in conv.state_dict().items()}
    torch.nn.utils.remove_weight_norm(conv)
    assert torch.allclose(layouts.fold_weight_norm(sd)["c.weight"], conv.weight, atol=1e-6)


def test_vocoder_chunk_limits(tmp_path):
    """Host logic of forward_batch's chunking: a chunk keeps every stage tensor below 2^31 elements and the workspace
    inside the budget; the configs of BASELINE.json run as one chunk."""
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    path = os.path.join(tmp_path, "b.pt")
    torch.save({"generator": factory.make_state_dict("bigvgan", 7)}, path)
    for streams, esz in (("f16", 2), ("f32", 4)):
        m = tb.BigVGAN(path, activation_dtype=streams)
        assert m._max_chunk(500) >= 64                       # config 2: 64 x 500 frames in one pass
        for frames in (100, 500, 1000, 4000, 20000):
            cap = m._max_chunk(frames)
            assert cap >= 1
            widest = max(m._stage_channels(i) * (frames * math.prod(m.upsample_scales[:i + 1]) + 16)
                         for i in range(len(m.upsample_scales)))
            assert cap * widest < (1 << 31) + widest          # per-stage element count of a chunk stays 32-bit addressable
        small = type(m).max_workspace_bytes
        m.max_workspace_bytes = 1 << 30
        assert m._max_chunk(1000) < tb.BigVGAN(path, activation_dtype=streams)._max_chunk(1000)
        m.max_workspace_bytes = small


def test_text_to_wave_batches_are_length_sorted_groups():
    from ims_toucan_prosody_variance_b200 import sharding
    lens = [5, 200, 37, 37, 120, 1, 64]
    batches = sharding.bucket_by_length(range(len(lens)), lens, max_batch=3)
    assert batches == [[1, 4, 6], [2, 3, 0], [5]]
