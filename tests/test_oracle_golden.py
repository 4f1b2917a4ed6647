"""Pins the oracle (oracle/dmel_oracle.py) to outputs of the unmodified reference (tests/golden, made by
scripts/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, oracle_kwargs, rel_err
from oracle import dmel_oracle as orc

NAMES = golden_names()


def test_fixtures_present():
    assert len(NAMES) >= 12


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("tag", ["lin", "log"])
def test_port_matches_reference_fp32(name, tag):
    g = load_golden(name)
    kw = oracle_kwargs(g["meta"])
    x = g["x"]
    dtype = torch.float64 if x.dtype == np.float64 else torch.float32
    # reference semantics for fp64 data: fp32 parameter, fp64 signal (datasets.py:33)
    xx = torch.as_tensor(x).clone().requires_grad_(True)
    lam = torch.tensor(float(g["lam"]), dtype=torch.float32, requires_grad=True)
    s = orc.port_forward(xx, lam, **kw)
    out = torch.log(s + orc.LOG_EPS) if tag == "log" else s
    (out * torch.as_tensor(g["G"])).sum().backward()
    # same library kernels underneath -> agreement far below the parity tolerance
    np.testing.assert_allclose(out.detach().numpy(), g[f"y_{tag}"], rtol=2e-5, atol=2e-6)
    assert rel_err(xx.grad.numpy(), g[f"dx_{tag}"]) < 2e-5
    assert abs(float(lam.grad) - float(g[f"dlam_{tag}"])) <= 2e-4 * abs(float(g[f"dlam_{tag}"])) + 1e-6
    assert dtype in (torch.float32, torch.float64)


@pytest.mark.parametrize("name", NAMES)
def test_batched_port_equals_loop_port(name):
    g = load_golden(name)
    kw = oracle_kwargs(g["meta"])
    x = torch.as_tensor(g["x"]).float()
    lam = torch.tensor(float(g["lam"]))
    a = orc.port_forward(x, lam, **kw)
    b = orc.port_forward_batched(x, lam, **kw)
    np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-4, atol=1e-5 * float(a.abs().max()))


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("tag", ["lin", "log"])
def test_closed_form_matches_reference_fp64(name, tag):
    """The explicit adjoint (what the CUDA backward implements) against the reference's fp64 autograd."""
    g = load_golden(name)
    kw = oracle_kwargs(g["meta"])
    out, dx, dlam = orc.closed_form(g["x"], float(g["lam"]), g["G"], use_log=(tag == "log"), **kw)
    # the reference rounds y to fp32 (models.py:36) even when run in fp64
    np.testing.assert_allclose(out, g[f"y64_{tag}"], rtol=3e-6, atol=3e-6 if tag == "log" else 1e-6 * np.abs(out).max())
    assert rel_err(dx, g[f"dx64_{tag}"]) < 2e-5  # the reference's fp64 run still rounds s to fp32 before the log
    assert abs(dlam - float(g[f"dlam64_{tag}"])) <= 2e-5 * abs(float(g[f"dlam64_{tag}"])) + 1e-9


def test_zero_mel_columns_are_log_eps():
    g = load_golden("mel_L128_zero_cols")
    y = g["y_log"]
    lin = g["y_lin"]
    dead = np.where(np.abs(lin).reshape(lin.shape[0], lin.shape[2], -1).max(axis=(0, 2)) == 0)[0]
    assert dead.size >= 1  # small n_fft -> some mel filters see no bin
    assert np.all(y[:, 0, dead, :] == np.float32(np.log(np.float32(1e-10))))


def test_window_length_rule():
    assert orc.next_power_of_2(0) == 2 and orc.next_power_of_2(1) == 1 and orc.next_power_of_2(2) == 2
    assert orc.next_power_of_2(3) == 4 and orc.next_power_of_2(512) == 512 and orc.next_power_of_2(513) == 1024
    assert orc.window_length(46.667, 8000, True) == (512, 512)
    assert orc.window_length(13.333, 8000, True) == (128, 128)
    assert orc.window_length(400.0, 8000, True) == (4096, 4096)
    assert orc.window_length(6.38, 128, False) == (128, 256)
