"""Structural drop-in check against the reference's own classifier heads (CPU, build container only: skipped where
/root/reference is absent, e.g. on the GPU box).  The heads are built once with the reference's layer and once with
ours swapped in; parameter names, shapes and state_dicts must be interchangeable (main.py:39, utils.py:269-270)."""
import os
import sys
import warnings

import pytest
import torch

import dmel_b200 as dm

REF = os.environ.get("DMEL_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "models.py")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref_models():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import models
    yield models
    sys.path.remove(REF)


def _build(models, name):
    lam = torch.tensor(46.667)
    if name == "mel_linear":
        return models.MelLinearNet(n_classes=10, init_lambd=lam, device="cpu", n_mels=64, sample_rate=8000, n_points=8000,
                                   hop_length=80, optimized=True, energy_normalize=True, normalize_window=False)
    if name == "mel_panns":
        return models.MelPANNsNet(n_classes=50, init_lambd=lam, device="cpu", n_mels=64, sample_rate=8000, n_points=40000,
                                  hop_length=80, optimized=True, energy_normalize=True, normalize_window=False)
    if name == "linear":
        return models.LinearNet(n_classes=3, init_lambd=torch.tensor(6.38), device="cpu", size=(129, 129), hop_length=1,
                                optimized=False, normalize_window=False)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["mel_linear", "mel_panns", "linear"])
def test_heads_accept_the_swapped_layer(ref_models, name, monkeypatch):
    ref_net = _build(ref_models, name)
    monkeypatch.setattr(ref_models, "MelSpectrogramLayer", dm.MelSpectrogramLayer)
    monkeypatch.setattr(ref_models, "SpectrogramLayer", dm.SpectrogramLayer)
    new_net = _build(ref_models, name)
    assert isinstance(new_net.spectrogram_layer, (dm.MelSpectrogramLayer, dm.SpectrogramLayer))
    ref_sd, new_sd = ref_net.state_dict(), new_net.state_dict()
    assert list(ref_sd.keys()) == list(new_sd.keys())
    assert all(ref_sd[k].shape == new_sd[k].shape and ref_sd[k].dtype == new_sd[k].dtype for k in ref_sd)
    assert "spectrogram_layer.lambd" in dict(new_net.named_parameters())          # main.py:39 lr group
    new_net.load_state_dict(ref_sd)                                               # utils.py:269-270 checkpoints
    assert float(new_net.spectrogram_layer.lambd) == float(ref_net.spectrogram_layer.lambd)
    new_net.spectrogram_layer.requires_grad_(False)                               # main.py:27 trainable=False
    assert not new_net.spectrogram_layer.lambd.requires_grad
    for attr in ("hop_length", "device", "optimized", "normalize_window"):
        assert getattr(new_net.spectrogram_layer, attr) == getattr(ref_net.spectrogram_layer, attr)
    if name != "linear":
        for attr in ("f_min", "f_max", "n_mels", "sample_rate", "n_freq", "n_time"):
            assert getattr(new_net.spectrogram_layer, attr) == getattr(ref_net.spectrogram_layer, attr)
