"""CPU: host-side pieces of the drop-in interface that need no GPU (language ids, PCM conversion, WAV writer, loud
failure without CUDA)."""
import os
import wave

import numpy as np
import pytest
import torch


def test_language_ids_match_reference_table():
    from ims_toucan_prosody_variance_b200 import interface
    # Preprocessing/TextFrontend.py:490-524
    assert interface.get_language_id("en").tolist() == [12]
    assert interface.get_language_id("de").tolist() == [1]
    assert interface.get_language_id("pt-br").tolist() == [17]
    assert len(interface.LANGUAGE_IDS) == 17 and sorted(interface.LANGUAGE_IDS.values()) == list(range(1, 18))
    with pytest.raises(KeyError):
        interface.get_language_id("xx")


def test_float2pcm_matches_reference_formula():
    from ims_toucan_prosody_variance_b200 import interface
    sig = torch.tensor([-1.5, -1.0, -0.5, 0.0, 0.25, 0.99999, 1.0, 2.0])
    ref = (sig.numpy() * 32768 + 0).clip(-32768, 32767).astype(np.int16)      # Utility/utils.py:20-33 for int16
    assert interface.float2pcm(sig).numpy().tolist() == ref.tolist()


def test_wav_writer_roundtrip(tmp_path):
    from ims_toucan_prosody_variance_b200 import interface
    pcm = (np.sin(np.arange(2400) / 10.0) * 20000).astype(np.int16)
    path = os.path.join(tmp_path, "x.wav")
    interface._write_pcm16(path, pcm, 24000)
    with wave.open(path, "rb") as f:
        assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (1, 2, 24000, 2400)
        assert np.frombuffer(f.readframes(2400), dtype=np.int16).tolist() == pcm.tolist()


def test_toucantts_refuses_cpu():
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    model = tb.ToucanTTS(weights=factory.make_state_dict("toucantts", 1234))
    with pytest.raises(tb._lib.EngineError):
        model.store_inverse_all()
    with pytest.raises(tb._lib.EngineError):
        model(factory.make_phoneme_tensor(5, 1), utterance_embedding=factory.make_utterance_embedding(1), lang_id=torch.tensor([12]))


def test_toucantts_state_dict_roundtrip_and_shared_wavenet_layers():
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    sd = factory.make_state_dict("toucantts", 3)
    model = tb.ToucanTTS(weights=sd)
    out = model.state_dict()
    assert list(out) == list(sd)
    for k in sd:
        assert torch.equal(out[k], sd[k]), k
    # WN in/res_skip layers are one storage within groups of 4 flow blocks (Glow.py:325-327)
    a = model.state_dict()["post_flow.flows.2.wn.in_layers.0.weight_v"]
    b = model.state_dict()["post_flow.flows.5.wn.in_layers.0.weight_v"]
    assert a.data_ptr() == b.data_ptr()


def test_pipeline_refuses_cpu_and_empty_list():
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    model = tb.ToucanTTS(weights=factory.make_state_dict("toucantts", 1234))
    eng = tb.TextToWave(model, None)
    assert eng.synthesize([], torch.zeros(64)) == []
    with pytest.raises(tb._lib.EngineError):
        eng.synthesize([factory.make_phoneme_tensor(5, 1)], torch.zeros(64))
