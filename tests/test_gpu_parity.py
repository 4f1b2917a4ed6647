"""Parity of the CUDA path (through the C ABI, via the drop-in layers / dmel op) against
  (1) tests/golden — outputs of the unmodified reference, (2) the CPU oracle on seeded inputs,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances are north_star's: log-mel rel 1e-4 / abs 1e-5; dL/dx, dL/dlambda rel 1e-3."""
import numpy as np
import pytest
import torch

import dmel_b200 as dm
from conftest import golden_names, load_golden, oracle_kwargs, rel_err
from oracle import dmel_oracle as orc

pytestmark = pytest.mark.gpu

Y_RTOL, Y_ATOL = 1e-4, 1e-5
G_RTOL = 1e-3
EPS = 1e-10


def op_kwargs(meta):
    kw = dict(hop_length=meta["hop"], optimized=meta["opt"], normalize_window=meta["norm"])
    if meta["kind"] == "mel":
        kw.update(n_mels=meta["M"], sample_rate=meta["sr"])
    return kw


def run_cuda(dev, x, lam, G, use_log, fused, need_dx=True, **kw):
    xx = torch.as_tensor(x).to(dev).clone().requires_grad_(need_dx)
    ll = torch.tensor(float(lam), device=dev, requires_grad=True)
    if use_log and fused:
        out = dm.dmel(xx, ll, log_eps=EPS, **kw)
    else:
        out = dm.dmel(xx, ll, **kw)
        if use_log:
            out = torch.log(out + EPS)
    (out * torch.as_tensor(G).to(dev)).sum().backward()
    torch.cuda.synchronize()
    return (out.detach().cpu().numpy(), xx.grad.detach().cpu().numpy() if need_dx else None, float(ll.grad))


def check_y(got, want, lin):
    if lin:
        # linear power: same relative bar, absolute floor scaled to the data (|y| spans decades)
        np.testing.assert_allclose(got, want, rtol=Y_RTOL, atol=Y_ATOL * float(np.abs(want).max()))
    else:
        np.testing.assert_allclose(got, want, rtol=Y_RTOL, atol=Y_ATOL)


def check_grads(dx, dlam, dx_ref, dlam_ref, dlam_ref64=None):
    assert rel_err(dx, dx_ref) < G_RTOL
    np.testing.assert_allclose(dx, dx_ref, rtol=G_RTOL, atol=G_RTOL * float(np.abs(dx_ref).max()))
    ok = abs(dlam - dlam_ref) <= G_RTOL * abs(dlam_ref)
    if not ok and dlam_ref64 is not None:  # the reference's own fp32 noise can exceed 1e-3 under cancellation
        ok = abs(dlam - dlam_ref64) <= G_RTOL * abs(dlam_ref64)
    assert ok, (dlam, dlam_ref, dlam_ref64)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("tag,fused", [("lin", False), ("log", True), ("log", False)])
def test_matches_reference_golden(cuda_device, name, tag, fused):
    g = load_golden(name)
    y, dx, dlam = run_cuda(cuda_device, g["x"], g["lam"], g["G"], tag == "log", fused, **op_kwargs(g["meta"]))
    assert y.dtype == np.float32 and y.shape == g[f"y_{tag}"].shape
    check_y(y, g[f"y_{tag}"], tag == "lin")
    check_grads(dx, dlam, g[f"dx_{tag}"], float(g[f"dlam_{tag}"]), float(g[f"dlam64_{tag}"]))
    # and against the fp64 run of the reference (tie-breaker: we should be at least as close to the truth)
    assert rel_err(dx, g[f"dx64_{tag}"]) < G_RTOL


@pytest.mark.parametrize("name", ["mel_L512_am", "mel_L2048_44k", "mel_L64_dft", "spec_tf_N64_small_lam"])
def test_lambda_only_backward_equals_full_backward(cuda_device, name):
    g = load_golden(name)
    kw = op_kwargs(g["meta"])
    _, _, full = run_cuda(cuda_device, g["x"], g["lam"], g["G"], True, True, need_dx=True, **kw)
    _, dx, only = run_cuda(cuda_device, g["x"], g["lam"], g["G"], True, True, need_dx=False, **kw)
    assert dx is None
    # lambda-only steps take the forward-mode (JVP) kernel: frequency-domain form of the same derivative
    assert abs(full - only) <= 2e-4 * abs(full)


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("tag", ["lin", "log"])
def test_jvp_lambda_gradient_matches_reference(cuda_device, name, tag, monkeypatch):
    """x without grad -> dL/dlambda comes from the JVP forward + dot product (tiled FFT sizes) or from the
    recomputing backward (big / direct-DFT sizes); both must match the reference, and each other."""
    g = load_golden(name)
    kw = op_kwargs(g["meta"])
    y, _, dlam = run_cuda(cuda_device, g["x"], g["lam"], g["G"], tag == "log", True, need_dx=False, **kw)
    check_y(y, g[f"y_{tag}"], tag == "lin")
    ref32, ref64 = float(g[f"dlam_{tag}"]), float(g[f"dlam64_{tag}"])
    assert abs(dlam - ref32) <= G_RTOL * abs(ref32) or abs(dlam - ref64) <= G_RTOL * abs(ref64)
    from importlib import import_module
    opmod = import_module("differentiable-mel-spectrogram_b200.op")
    monkeypatch.setattr(opmod, "USE_JVP", False)
    _, _, dlam_rc = run_cuda(cuda_device, g["x"], g["lam"], g["G"], tag == "log", True, need_dx=False, **kw)
    assert abs(dlam - dlam_rc) <= 2e-4 * abs(dlam_rc) + 1e-7


def test_jvp_backward_can_run_twice_and_on_baseline_shape(cuda_device):
    dev = cuda_device
    g = torch.Generator().manual_seed(21)
    x = torch.randn(256, 8000, generator=g)
    G = torch.randn(256, 1, 64, 101, generator=g)
    lam = 8000 * 0.035 / 6
    want = orc.port_forward_backward(x, lam, G, use_log=True, batched=True, hop=80, optimized=True, n_mels=64, sample_rate=8000)
    ll = torch.tensor(lam, device=dev, requires_grad=True)
    y = dm.dmel(x.to(dev), ll, hop_length=80, optimized=True, n_mels=64, sample_rate=8000, log_eps=EPS)
    (y * G.to(dev)).sum().backward(retain_graph=True)
    first = float(ll.grad)
    ll.grad = None
    (y * G.to(dev)).sum().backward()          # the fp64 accumulator in the workspace resets itself
    assert float(ll.grad) == first
    assert abs(first - want[2]) <= G_RTOL * abs(want[2])
    check_y(y.detach().cpu().numpy(), want[0], False)


def _oracle_case(dev, B, N, sr, hop, M, lam, seed, use_log=True, norm=False, optimized=True, x_dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, generator=g).to(x_dtype)
    kw_o = dict(hop=hop, optimized=optimized, norm=norm, n_mels=M, sample_rate=sr)
    kw_c = dict(hop_length=hop, optimized=optimized, normalize_window=norm, n_mels=M, sample_rate=sr)
    with torch.no_grad():
        shape = orc.port_forward_batched(x[:1].float(), torch.tensor(lam), **kw_o).shape
    G = torch.randn((B,) + tuple(shape[1:]), generator=g)
    want = orc.port_forward_backward(x.float(), lam, G, use_log=use_log, batched=True, **kw_o)
    got = run_cuda(dev, x, lam, G, use_log, True, **kw_c)
    return got, want


# BASELINE.json configs at sizes the oracle finishes in seconds (full AM batch; ESC / large-batch clips at reduced B)
@pytest.mark.parametrize("B,N,sr,hop,M,lam", [
    (256, 8000, 8000, 80, 64, 8000 * 0.035 / 6),        # configs[1] audio_mnist, full size, n_fft 512
    (4, 220500, 44100, 441, 128, 44100 * 0.035 / 6),    # configs[2] esc50 clip shape, n_fft 2048
    (4, 220500, 44100, 441, 128, 44100 * 0.010 / 6),    # esc50 lambda sweep: 10 ms -> n_fft 512
    (16, 8000, 8000, 80, 64, 0.75 * 256 / 6),           # configs[3] window sweep n_fft 256
    (16, 8000, 8000, 80, 64, 0.75 * 1024 / 6),          # n_fft 1024
    (16, 8000, 8000, 80, 64, 0.75 * 2048 / 6),          # n_fft 2048
    (16, 8000, 8000, 80, 64, 0.75 * 4096 / 6),          # n_fft 4096
    (16, 8000, 8000, 80, 64, 8000 * 0.010 / 6),         # reference grid 10 ms -> n_fft 128
    (16, 8000, 8000, 80, 64, 8000 * 0.300 / 6),         # reference grid 300 ms -> n_fft 4096
    (3, 160000, 16000, 160, 128, 16000 * 0.035 / 6),    # configs[4] large-batch clip shape, n_fft 1024
    (2, 220500, 44100, 441, 128, 44100 * 0.300 / 6),    # esc50 lambda sweep: 300 ms -> n_fft 16384 (big-FFT path)
    (2, 160000, 16000, 160, 128, 16000 * 0.300 / 6),    # 300 ms at 16 kHz -> n_fft 8192 (big-FFT path)
])
def test_matches_oracle_on_baseline_shapes(cuda_device, B, N, sr, hop, M, lam):
    (y, dx, dlam), (y0, dx0, dlam0) = _oracle_case(cuda_device, B, N, sr, hop, M, lam, seed=1234)
    check_y(y, y0, False)
    check_grads(dx, dlam, dx0, dlam0)


def test_time_frequency_config_spectrogram_layer(cuda_device):
    """configs[0]: SpectrogramLayer, optimized=False, hop 1, N=256 -> n_fft 512, out (B,1,257,257); fp64 inputs."""
    for N, lam in [(256, 6.38), (128, 0.2 * 6.38), (128, 5 * 6.38)]:
        g = torch.Generator().manual_seed(N)
        x = torch.randn(8, N, generator=g, dtype=torch.float64)
        G = torch.randn(8, 1, N + 1, N + 1, generator=g)
        want = orc.port_forward_backward(x.float(), lam, G, use_log=False, batched=True, hop=1, optimized=False)
        layer = dm.SpectrogramLayer(torch.tensor(lam), device="cuda:0", optimized=False, size=(N + 1, N + 1), hop_length=1).to(cuda_device)
        xx = x.to(cuda_device).requires_grad_(True)
        s = layer(xx)
        assert s.shape == (8, 1, N + 1, N + 1) and s.dtype == torch.float32 and s.is_contiguous()
        (s * G.to(cuda_device)).sum().backward()
        check_y(s.detach().cpu().numpy(), want[0], True)
        check_grads(xx.grad.cpu().numpy(), float(layer.lambd.grad), want[1], want[2])
        assert xx.grad.dtype == torch.float64  # gradient comes back in the caller's dtype


def test_full_size_properties_esc50_and_large_batch(cuda_device):
    """At BASELINE's full sizes the oracle is too slow; use exact, size-independent properties instead."""
    dev = cuda_device
    for (B, N, sr, hop, M, lam) in [(128, 220500, 44100, 441, 128, 257.25), (256, 160000, 16000, 160, 128, 93.33)]:
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, N, generator=g).to(dev)
        ll = torch.tensor(lam, device=dev, requires_grad=True)
        kw = dict(hop_length=hop, optimized=True, n_mels=M, sample_rate=sr)
        y = dm.dmel(x, ll, **kw).detach()
        assert y.shape == (B, 1, M, N // hop + 1) and y.is_contiguous() and torch.isfinite(y).all()
        # (a) clips are independent: permuting the batch permutes the output bit-exactly
        perm = torch.randperm(B, generator=g).to(dev)
        assert torch.equal(dm.dmel(x[perm], ll, **kw).detach(), y[perm])
        # (b) sharding the batch (what the multi-GPU path does) changes nothing, bit for bit
        h = B // 2
        assert torch.equal(torch.cat([dm.dmel(x[:h], ll, **kw), dm.dmel(x[h:], ll, **kw)]).detach(), y)
        # (c) power is quadratic: scaling by 2 scales every output by exactly 4 in fp32; DC offset is removed
        assert torch.equal(dm.dmel(2.0 * x, ll, **kw).detach(), 4.0 * y)
        y_dc = dm.dmel(x + 3.0, ll, **kw)
        assert rel_err(y_dc.detach().cpu().numpy(), y.detach().cpu().numpy()) < 1e-4
        # (d) d/dlambda accumulated over shards equals the whole-batch gradient; dx rows have zero mean
        G = torch.randn(y.shape, generator=torch.Generator().manual_seed(9)).to(dev)
        xx = x.clone().requires_grad_(True)
        yl = dm.dmel(xx, ll, log_eps=EPS, **kw)
        (yl * G).sum().backward()
        whole, dx = float(ll.grad), xx.grad
        ll.grad = None
        for sl in (slice(0, h), slice(h, B)):
            (dm.dmel(x[sl], ll, log_eps=EPS, **kw) * G[sl]).sum().backward()
        # (the shards take the forward-mode kernel since their x carries no grad: same derivative, frequency-domain form)
        assert abs(float(ll.grad) - whole) <= 2e-4 * abs(whole)
        assert float(dx.mean(dim=1).abs().max()) <= 1e-4 * float(dx.abs().max())
        # (e) a few clips against the oracle
        sub = [0, B // 2, B - 1]
        want = orc.port_forward_backward(x[sub].cpu(), lam, G[sub].cpu(), use_log=True, batched=True, hop=hop,
                                         optimized=True, n_mels=M, sample_rate=sr)
        check_y(yl[sub].detach().cpu().numpy(), want[0], False)
        assert rel_err(dx[sub].cpu().numpy(), want[1]) < G_RTOL
        ll.grad = None
        del x, y, G, xx, yl, dx
        torch.cuda.empty_cache()


@pytest.mark.parametrize("B,N,hop,lam,M", [
    (1, 50, 80, 46.667, 8),      # N < hop: a single frame, clip shorter than half a window
    (2, 79, 80, 20.0, 8),        # T = 1
    (3, 161, 80, 46.667, 16),    # T = 3 (odd): last pair has one real frame
    (2, 1000, 600, 13.333, 8),   # hop > n_fft: frames leave gaps in x (dx there is only -mean)
    (5, 4001, 1, 23.0, 12),      # hop 1, many tiles
    (2, 2500, 33, 90.0, 24),     # n_fft 1024 with an odd hop
])
def test_edge_geometries_against_oracle(cuda_device, B, N, hop, lam, M):
    (y, dx, dlam), (y0, dx0, dlam0) = _oracle_case(cuda_device, B, N, 8000, hop, M, lam, seed=N + hop)
    check_y(y, y0, False)
    check_grads(dx, dlam, dx0, dlam0)


def test_strided_rows_fp16_input_and_negative_lambda(cuda_device):
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    big = torch.randn(4, 3000, generator=g)
    x = big[:, 100:1700]                                   # row stride 3000, offset start
    G = torch.randn(4, 1, 32, 21, generator=g)
    kw_o = dict(hop=80, optimized=True, n_mels=32, sample_rate=8000)
    want = orc.port_forward_backward(x.contiguous(), -46.667, G, use_log=True, batched=True, **kw_o)
    xb = big.to(dev)
    xs = xb[:, 100:1700].requires_grad_(True)
    ll = torch.tensor(-46.667, device=dev, requires_grad=True)
    y = dm.dmel(xs, ll, hop_length=80, optimized=True, n_mels=32, sample_rate=8000, log_eps=EPS)
    (y * G.to(dev)).sum().backward()
    check_y(y.detach().cpu().numpy(), want[0], False)
    check_grads(xs.grad.cpu().numpy(), float(ll.grad), want[1], want[2])
    assert want[2] * float(ll.grad) > 0
    # half-precision storage is converted on entry
    xh = x.to(torch.float16)
    yh = dm.dmel(xh.to(dev), ll, hop_length=80, optimized=True, n_mels=32, sample_rate=8000)
    w = orc.port_forward_batched(xh.float(), torch.tensor(46.667), **kw_o)
    check_y(yh.detach().cpu().numpy(), w.numpy(), True)


@pytest.mark.parametrize("offset,N", [(1, 1600), (2, 1603), (3, 4001), (101, 1599),
                                      (1, 70001), (0, 98306)])   # > 32768 points: the mean is summed by a CTA cluster
def test_unaligned_input_rows_take_the_scalar_copy_path(cuda_device, offset, N):
    """x rows that are only 4-byte aligned (a column slice of a bigger buffer): the 16-byte async copies must not be used."""
    dev = cuda_device
    g = torch.Generator().manual_seed(offset + N)
    big = torch.randn(3, N + 2 * offset + 5, generator=g)
    x = big[:, offset:offset + N]
    T = N // 80 + 1
    G = torch.randn(3, 1, 32, T, generator=g)
    kw_o = dict(hop=80, optimized=True, n_mels=32, sample_rate=8000)
    want = orc.port_forward_backward(x.contiguous(), 46.667, G, use_log=True, batched=True, **kw_o)
    xs = big.to(dev)[:, offset:offset + N].requires_grad_(True)
    assert xs.data_ptr() % 16 != 0 or xs.stride(0) % 4 != 0
    ll = torch.tensor(46.667, device=dev, requires_grad=True)
    y = dm.dmel(xs, ll, hop_length=80, optimized=True, n_mels=32, sample_rate=8000, log_eps=EPS)
    (y * G.to(dev)).sum().backward()
    check_y(y.detach().cpu().numpy(), want[0], False)
    check_grads(xs.grad.cpu().numpy(), float(ll.grad), want[1], want[2])


def test_zero_padded_tone_bursts_low_energy_bins(cuda_device):
    """Real clips are zero-padded speech / events (datasets.py:175): most frames are silence plus the DC step left by the
    mean removal, and many mel bins sit 100+ dB below the peak.  There FFT rounding differs by algorithm, so the linear
    mel is compared with an absolute floor scaled to the clip's peak, and the log output only where the bin carries energy."""
    dev = cuda_device
    N, sr, hop, M, lam = 8000, 8000, 80, 64, 46.667
    t = torch.arange(N) / sr
    x = torch.zeros(4, N)
    x[0, 1000:3000] = torch.sin(2 * torch.pi * 440 * t[1000:3000])
    x[1, 500:900] = torch.sin(2 * torch.pi * 1500 * t[500:900]) * torch.hann_window(400)
    x[2, 4000:4100] = 1.0                                   # a click
    x[3, :] = 1e-3 * torch.randn(N, generator=torch.Generator().manual_seed(2))
    G = torch.randn(4, 1, M, N // hop + 1, generator=torch.Generator().manual_seed(3))
    kw_o = dict(hop=hop, optimized=True, n_mels=M, sample_rate=sr)
    lin64 = orc.port_forward_batched(x.double(), torch.tensor(lam, dtype=torch.float64), out_dtype=torch.float64, **kw_o).numpy()
    ll = torch.tensor(lam, device=dev, requires_grad=True)
    xin = x.to(dev).requires_grad_(True)
    lin = dm.dmel(xin, ll, hop_length=hop, optimized=True, n_mels=M, sample_rate=sr)
    got = lin.detach().cpu().numpy().astype(np.float64)
    for b in range(4):
        peak = lin64[b].max()
        np.testing.assert_allclose(got[b], lin64[b], rtol=1e-4, atol=1e-6 * peak)
    ylog = dm.dmel(xin, ll, hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=EPS).detach().cpu().numpy()
    strong = lin64 > 1e-5 * lin64.reshape(4, -1).max(axis=1)[:, None, None, None]
    np.testing.assert_allclose(ylog[strong], np.log(lin64[strong] + EPS), rtol=Y_RTOL, atol=Y_ATOL)
    assert np.isfinite(ylog).all() and ylog.min() >= np.log(EPS) - 1e-3
    # gradients through the linear output against the fp64 closed form
    _, dx64, dl64 = orc.closed_form(x.numpy(), lam, G.numpy(), use_log=False, **kw_o)
    (lin * G.to(dev)).sum().backward()
    assert rel_err(xin.grad.cpu().numpy(), dx64) < G_RTOL
    assert abs(float(ll.grad) - dl64) <= G_RTOL * abs(dl64)


def test_all_zero_input_hits_the_log_floor_exactly(cuda_device):
    layer = dm.MelSpectrogramLayer(torch.tensor(46.667), n_mels=64, n_points=1600, sample_rate=8000, hop_length=80,
                                   device="cuda:0", optimized=True, log_eps=EPS).to(cuda_device)
    y = layer(torch.zeros(2, 1600, device=cuda_device) + 5.0)   # constant clip: mean removal leaves silence
    assert torch.all(y == torch.log(torch.tensor(EPS, device=cuda_device)))


def test_drop_in_module_contract(cuda_device):
    """What the reference's heads do with the layer (models.py:59-78, main.py:27,36-48, utils.py:269)."""
    dev = cuda_device

    class MelLinear(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.spectrogram_layer = dm.MelSpectrogramLayer(torch.tensor(46.667), n_mels=64, n_points=8000, sample_rate=8000,
                                                            hop_length=80, device="cuda:0", optimized=True)
            self.fc = torch.nn.Linear(64 * 101, 10)

        def forward(self, x):
            s = self.spectrogram_layer(x)
            s = torch.log(s + 1e-10)
            return self.fc(s.view(-1, 64 * 101)), s

    net = MelLinear().to(dev)
    names = [n for n, _ in net.named_parameters()]
    assert "spectrogram_layer.lambd" in names
    x = torch.randn(8, 8000, device=dev)
    logits, s = net(x)
    assert s.shape == (8, 1, 64, 101) and logits.shape == (8, 10)
    torch.nn.functional.cross_entropy(logits, torch.arange(8, device=dev) % 10).backward()
    assert net.spectrogram_layer.lambd.grad is not None and torch.isfinite(net.spectrogram_layer.lambd.grad)
    assert isinstance(net.spectrogram_layer.lambd.item(), float)
    sd = net.state_dict()
    net2 = MelLinear().to(dev)
    net2.load_state_dict(sd)
    assert torch.equal(net2(x)[1], s)
    # frozen layer (trainable=False, main.py:27): no graph through the op
    net.spectrogram_layer.requires_grad_(False)
    assert not net(x)[1].requires_grad
    # lambda moved by the optimiser -> n_fft follows it (512 -> 1024)
    with torch.no_grad():
        net.spectrogram_layer.lambd.fill_(90.0)
    assert net(x)[1].shape == (8, 1, 64, 101)


@pytest.mark.parametrize("lam,M,sr,hop,N", [(46.667, 64, 8000, 80, 8000), (93.33, 128, 16000, 160, 16000), (257.25, 128, 44100, 441, 22050),
                                            (800.0, 64, 16000, 160, 16000)])      # last: n_fft 8192, the big-FFT path
def test_time_major_layout_is_the_exact_transpose(cuda_device, lam, M, sr, hop, N):
    """MelPANNsNet transposes the layer output to (B,1,T,M) (models.py:162); the kernels can emit it."""
    dev = cuda_device
    g = torch.Generator().manual_seed(int(lam))
    x = torch.randn(6, N, generator=g).to(dev)
    T = N // hop + 1
    G = torch.randn(6, 1, M, T, generator=g).to(dev)
    kw = dict(hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=EPS)
    for need_dx in (True, False):                      # recomputing backward / forward-mode (JVP) path
        res = []
        for tm in (False, True):
            xin = x.clone().requires_grad_(need_dx)
            ll = torch.tensor(lam, device=dev, requires_grad=True)
            y = dm.dmel(xin, ll, time_major=tm, **kw)
            assert y.is_contiguous() and y.shape == ((6, 1, T, M) if tm else (6, 1, M, T))
            (y * (G.transpose(2, 3) if tm else G)).sum().backward()
            res.append((y.detach(), xin.grad, float(ll.grad)))
        assert torch.equal(res[1][0], res[0][0].transpose(2, 3))
        if need_dx:
            assert rel_err(res[1][1].cpu().numpy(), res[0][1].cpu().numpy()) < 1e-6
        assert abs(res[1][2] - res[0][2]) <= 1e-6 * abs(res[0][2])
    layer = dm.MelSpectrogramLayer(torch.tensor(lam), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop, device="cuda:0",
                                   optimized=True, log_eps=EPS, time_major=True).to(dev)
    assert layer(x).shape == (6, 1, T, M)


def test_fused_energy_sum_equals_output_sum(cuda_device):
    """train.py:58 copies every spectrogram batch to the host to add it up; the kernel leaves the sum."""
    dev = cuda_device
    x = torch.randn(32, 8000, generator=torch.Generator().manual_seed(11)).to(dev)
    for log_eps in (None, EPS):
        layer = dm.MelSpectrogramLayer(torch.tensor(46.667), n_mels=64, n_points=8000, sample_rate=8000, hop_length=80,
                                       device="cuda:0", optimized=True, log_eps=log_eps).to(dev)
        y = layer(x)
        want = float(y.detach().double().sum())
        assert layer.last_energy.dtype == torch.float64 and layer.last_energy.dim() == 0
        assert abs(float(layer.last_energy) - want) <= 1e-6 * abs(want)
    spec = dm.SpectrogramLayer(torch.tensor(6.38), device="cuda:0", optimized=False, size=(129, 129), hop_length=1).to(dev)
    s = spec(torch.randn(16, 128, generator=torch.Generator().manual_seed(12)).to(dev))
    assert abs(float(spec.last_energy) - float(s.detach().double().sum())) <= 1e-6 * float(s.detach().double().sum())


def test_time_frequency_function_mirror(cuda_device):
    dev = cuda_device
    x = torch.randn(1000, generator=torch.Generator().manual_seed(5))
    want = orc._spectrogram_1d(x, torch.tensor(23.0), True, 50, False)
    got, w = dm.differentiable_spectrogram(x.to(dev), torch.tensor(23.0, device=dev), optimized=True, hop_length=50,
                                           return_window=True)
    check_y(got.cpu().numpy(), want.numpy(), True)
    np.testing.assert_allclose(w.cpu().numpy(), orc.gaussian_window(torch.tensor(23.0), 256, False).numpy(), rtol=1e-6)


def _random_eval_bn(M, seed):
    g = torch.Generator().manual_seed(seed)
    bn = torch.nn.BatchNorm2d(M).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(M, generator=g) * 1.5 + 0.25)
        bn.weight[::7] *= -1.0                                  # negative gains are legal
        bn.bias.copy_(torch.randn(M, generator=g))
        bn.running_mean.copy_(torch.randn(M, generator=g) * 3.0 - 2.0)
        bn.running_var.copy_(torch.rand(M, generator=g) * 4.0 + 0.5)
    return bn


@pytest.mark.parametrize("B,N,sr,hop,M,lam,optimized,use_log,tm", [
    (8, 8000, 8000, 80, 64, 46.667, True, True, False),         # tiled FFT path, n_fft 512
    (8, 8000, 8000, 80, 64, 46.667, True, True, True),          #   ... emitting (B,1,T,M) as Cnn6 takes it
    (8, 8000, 8000, 80, 64, 46.667, True, False, False),        #   ... linear mel power
    (3, 22050, 44100, 441, 128, 257.25, True, True, True),      # esc50 clip shape (short), n_fft 2048
    (2, 32000, 16000, 160, 128, 16000 * 0.300 / 6, True, True, False),   # big-FFT path, n_fft 8192
    (4, 100, 8000, 10, 16, 5.0, False, True, False),            # direct-DFT path, n_fft 200
])
def test_folded_eval_batchnorm_matches_batchnorm_after_the_layer(cuda_device, B, N, sr, hop, M, lam, optimized, use_log, tm):
    """Cnn6 starts with BatchNorm2d over mel bins on the transposed log-mel (panns.py:169-172).  In
    eval mode that is a per-mel affine map; folded into the epilogue it must equal the real BatchNorm2d applied to
    the oracle's output, with the same gradients to x and lambd."""
    dev = cuda_device
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, N, generator=g)
    bn = _random_eval_bn(M, 7)
    kw_o = dict(hop=hop, optimized=optimized, norm=False, n_mels=M, sample_rate=sr)
    with torch.no_grad():
        y0 = orc.port_forward_batched(x, torch.tensor(lam), **kw_o)             # (B,1,M,T) linear mel power
        if use_log:
            y0 = torch.log(y0 + EPS)
        want_y = bn(y0.transpose(1, 2)).transpose(1, 2)                         # channels = mel bins
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    G = torch.randn(want_y.shape, generator=g)
    # BatchNorm in eval mode is affine per mel: the gradient reaching the layer is G * scale[m]
    _, dx0, dlam0 = orc.port_forward_backward(x, lam, G * scale.view(1, 1, M, 1), use_log=use_log, batched=True, **kw_o)

    layer = dm.MelSpectrogramLayer(torch.tensor(lam), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop, device="cuda:0",
                                   optimized=optimized, log_eps=EPS if use_log else None, time_major=tm).to(dev)
    layer.fold_batchnorm(bn)
    assert list(layer.state_dict()) == ["lambd"]
    Gd = (G.transpose(2, 3) if tm else G).contiguous().to(dev)
    for need_dx in (True, False):                     # recomputing backward, then the forward-mode (JVP) path
        layer.lambd.grad = None
        xin = x.to(dev).requires_grad_(need_dx)
        y = layer(xin)
        (y * Gd).sum().backward()
        got = y.detach().transpose(2, 3) if tm else y.detach()
        span = float(want_y.abs().max())
        np.testing.assert_allclose(got.cpu().numpy(), want_y.numpy(), rtol=Y_RTOL,
                                   atol=Y_ATOL * (max(1.0, span) if use_log else span))
        assert abs(float(layer.last_energy) - float(y.double().sum())) <= 1e-6 * float(y.double().abs().sum())
        dlam = float(layer.lambd.grad)
        assert abs(dlam - dlam0) <= G_RTOL * abs(dlam0), (dlam, dlam0)
        if need_dx:
            assert rel_err(xin.grad.cpu().numpy(), dx0) < G_RTOL
    # removing the map restores the plain output
    layer.set_output_affine(None)
    with torch.no_grad():
        y_plain = layer(x.to(dev))
    y_plain = y_plain.transpose(2, 3) if tm else y_plain
    np.testing.assert_allclose(y_plain.cpu().numpy(), y0.numpy(), rtol=Y_RTOL,
                               atol=Y_ATOL * (1.0 if use_log else float(y0.abs().max())))


def test_layer_forward_backward_can_be_captured_in_a_cuda_graph(cuda_device):
    """The op allocates through torch's allocator, never synchronises and reads lambd from the host mirror while the
    parameter is unchanged, so a whole training-style step of the layer can live in a CUDA graph."""
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    layer = dm.MelSpectrogramLayer(torch.tensor(46.667), n_mels=64, n_points=8000, sample_rate=8000, hop_length=80,
                                   device="cuda:0", optimized=True, log_eps=EPS).to(dev)
    x_static = torch.randn(16, 8000, generator=g).to(dev)
    G = torch.randn(16, 1, 64, 101, generator=g).to(dev)

    def step():
        layer.lambd.grad = None
        y = layer(x_static)
        (y * G).sum().backward()
        return y

    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(3):                        # warm-up on a side stream, as torch.cuda.graph wants
            step()
    torch.cuda.current_stream(dev).wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    layer.lambd.grad = None
    with torch.cuda.graph(graph):
        y_static = layer(x_static)
        (y_static * G).sum().backward()
    grad_static = layer.lambd.grad
    for seed in (11, 12):
        x_new = torch.randn(16, 8000, generator=torch.Generator().manual_seed(seed)).to(dev)
        x_static.copy_(x_new)
        graph.replay()
        torch.cuda.synchronize()
        got_y, got_g = y_static.clone(), float(grad_static)
        want_y = step()
        assert torch.equal(got_y, want_y.detach())
        assert abs(got_g - float(layer.lambd.grad)) <= 1e-6 * abs(got_g)


def test_per_clip_mean_is_batch_invariant_and_exact_for_long_clips(cuda_device):
    """models.py:38 removes each clip's mean.  Clips longer than 32768 points are summed by a cluster of CTAs whose
    split depends on N only: the same clip gives bit-identical output alone or inside a batch, and a large DC offset
    (1e3 x the signal) is removed to fp32 accuracy (the sum is carried in fp64)."""
    dev = cuda_device
    g = torch.Generator().manual_seed(17)
    N = 220500
    x = torch.randn(5, N, generator=g) + torch.tensor([0.0, 1000.0, -1000.0, 3.0, 0.5]).view(5, 1)
    kw = dict(hop_length=441, optimized=True, n_mels=128, sample_rate=44100, log_eps=EPS)
    ll = torch.tensor(257.25, device=dev)
    xd = x.to(dev)
    with torch.no_grad():
        y_all = dm.dmel(xd, ll, **kw)
        for i in range(5):
            assert torch.equal(dm.dmel(xd[i:i + 1], ll, **kw), y_all[i:i + 1])
        # same clips with the offset removed in fp64 on the host
        x0 = (x.double() - x.double().mean(dim=1, keepdim=True)).float().to(dev)
        y_ref = dm.dmel(x0, ll, **kw)
    # the offset clips lost ~10 bits of signal mantissa to the DC before reaching the kernel; compare the rest
    np.testing.assert_allclose(y_all[[0, 3, 4]].cpu().numpy(), y_ref[[0, 3, 4]].cpu().numpy(), rtol=Y_RTOL, atol=2e-4)
    np.testing.assert_allclose(y_all[[1, 2]].cpu().numpy(), y_ref[[1, 2]].cpu().numpy(), rtol=0, atol=0.05)


def _fuzz_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < n:
        n_fft = int(2 ** rng.integers(5, 13))                    # 32 .. 4096 (direct-DFT below 128, FFT above)
        lam = float(rng.uniform(0.52, 0.99) * n_fft / 6.0)       # inside the bracket that derives n_fft
        hop = int(rng.choice([1, 3, 17, 80, 160, 441, 700]))
        N = int(rng.integers(max(hop, 40), 6000))
        T = N // hop + 1
        if T * n_fft > 3_000_000:                                # keep the CPU oracle in the seconds
            continue
        M = int(rng.choice([0, 8, 23, 64, 80]))
        if M and M > n_fft // 4:
            continue
        cases.append((int(rng.integers(1, 6)), N, hop, lam, M, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), n_fft))
    return cases


@pytest.mark.parametrize("B,N,hop,lam,M,norm,use_log,n_fft", _fuzz_cases(28, 2024))
def test_random_geometries_against_oracle(cuda_device, B, N, hop, lam, M, norm, use_log, n_fft):
    """Seeded sweep over odd shapes (ragged last tiles, hop > n_fft, hop 1, spectrogram and mel mode, normalised window,
    linear and log output) against the CPU oracle, forward and both gradients."""
    assert orc.window_length(lam, N, True)[1] == n_fft
    dev = cuda_device
    g = torch.Generator().manual_seed(N * 7 + hop)
    x = torch.randn(B, N, generator=g)
    kw_o = dict(hop=hop, optimized=True, norm=norm)
    kw_c = dict(hop_length=hop, optimized=True, normalize_window=norm)
    if M:
        kw_o.update(n_mels=M, sample_rate=8000)
        kw_c.update(n_mels=M, sample_rate=8000)
    rows = M if M else n_fft // 2 + 1
    G = torch.randn(B, 1, rows, N // hop + 1, generator=g)
    want = orc.port_forward_backward(x, lam, G, use_log=use_log, batched=True, **kw_o)
    got = run_cuda(dev, x, lam, G, use_log, True, **kw_c)
    check_y(got[0], want[0], not use_log)
    # the oracle's own fp32 gradient noise grows with cancellation; fall back to its fp64 run when it is the outlier
    try:
        check_grads(got[1], got[2], want[1], want[2])
    except AssertionError as first:
        # logged (pytest -rA / captured output), so that a kernel drifting towards the bar cannot hide in this branch
        print(f"[fp64 fallback] B={B} N={N} hop={hop} lam={lam} M={M} norm={norm} log={use_log} n_fft={n_fft}: "
              f"fp32 oracle comparison failed ({str(first).splitlines()[0][:120]}); comparing with the fp64 oracle")
        FP64_FALLBACKS.append((B, N, hop, lam, M, norm, use_log, n_fft))
        want64 = orc.port_forward_backward(x, lam, G, use_log=use_log, batched=True, dtype=torch.float64, **kw_o)
        check_grads(got[1], got[2], np.asarray(want64[1], dtype=np.float64), float(want64[2]))


FP64_FALLBACKS = []


def test_fp64_fallback_stays_rare():
    """Runs after the sweep above (file order): at most a handful of the 28 seeded geometries may need the fp64 oracle
    (round 1: 2).  A kernel whose fp32 error grows shows up here first."""
    assert len(FP64_FALLBACKS) <= 4, FP64_FALLBACKS


# ------------------------------------------------------------------------------------------ round 2 additions
@pytest.mark.parametrize("dtype", [torch.int16, torch.float16])
@pytest.mark.parametrize("B,N,sr,hop,M,lam", [
    (5, 8000, 8000, 80, 64, 8000 * 0.035 / 6),          # n_fft 512
    (2, 16001, 16000, 160, 128, 16000 * 0.035 / 6),     # n_fft 1024, odd length: rows are not 8-byte aligned
    (2, 22050, 44100, 441, 128, 44100 * 0.035 / 6),     # n_fft 2048
])
def test_sixteen_bit_samples_are_widened_in_kernel(cuda_device, dtype, B, N, sr, hop, M, lam):
    """int16 PCM / fp16 clips go to the kernels as they are (DMEL_FLAG_X_I16 / _F16, dmel.h) and must
    give what the oracle gives on x.to(float32) — forward, the forward-mode lambda gradient and the recomputing one."""
    dev = cuda_device
    lib = dm.load()
    assert lib.dmel_x16_supported(orc.window_length(lam, N, True)[1])
    g = torch.Generator().manual_seed(11)
    xf = torch.randn(B, N, generator=g)
    x16 = (xf * 3000.0).round().clamp(-32768, 32767).to(torch.int16) if dtype == torch.int16 else xf.to(torch.float16)
    G = torch.randn(B, 1, M, N // hop + 1, generator=g)
    want = orc.port_forward_backward(x16.to(torch.float32), lam, G, use_log=True, batched=True, hop=hop, optimized=True,
                                     n_mels=M, sample_rate=sr)
    for use_jvp in (True, False):
        dm.op.USE_JVP = use_jvp
        try:
            ll = torch.tensor(float(lam), device=dev, requires_grad=True)
            xd = x16.to(dev)
            y = dm.dmel(xd, ll, hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=EPS)
            (y * G.to(dev)).sum().backward()
        finally:
            dm.op.USE_JVP = True
        check_y(y.detach().cpu().numpy(), want[0], False)
        assert abs(float(ll.grad) - want[2]) <= G_RTOL * abs(want[2]), (use_jvp, float(ll.grad), want[2])
    # and the same numbers as the fp32 path on the widened tensor, to rounding of the mean
    y32 = dm.dmel(x16.to(dev).to(torch.float32), ll, hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=EPS)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y32.detach().cpu().numpy(), rtol=Y_RTOL, atol=Y_ATOL)


@pytest.mark.parametrize("N,sr,hop,M,lam", [(8000, 8000, 80, 64, 46.667), (16000, 16000, 160, 128, 93.33),
                                            (4000, 8000, 40, 64, 30.0)])
def test_segment_projection_equals_mel_major_gather(cuda_device, N, sr, hop, M, lam):
    """dmel.h dmel_build_segment_tables: with the bank's segment tables the forward projects segment by segment, without
    them mel by mel — same weights, same sums up to fp32 summation order."""
    import ctypes
    from importlib import import_module
    mb = import_module("differentiable-mel-spectrogram_b200.melbank")
    dev = cuda_device
    lib = dm.load()
    geo = dm.resolve_geometry(3, N, hop, lam, True, M)
    bank = mb.device_bank(geo.n_freqs, 0.0, sr // 2, M, sr, dev)
    assert bank.struct.seg_tables, "triangular HTK banks of these sizes have usable segment tables"
    x = torch.randn(3, N, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    ws = torch.empty(lib.dmel_workspace_bytes(3, geo.n_fft), dtype=torch.uint8, device=dev)
    outs = []
    for with_tables in (True, False):
        s = mb.MelBankStruct()
        ctypes.memmove(ctypes.byref(s), ctypes.byref(bank.struct), ctypes.sizeof(mb.MelBankStruct))
        if not with_tables:
            s.seg_tables = None
        y = torch.empty(3, 1, M, geo.n_frames, device=dev)
        rc = lib.dmel_forward_f32(x.data_ptr(), 3, N, N, lam, geo.win_length, geo.n_fft, hop, M, ctypes.byref(s), EPS, 0,
                                  y.data_ptr(), ws.data_ptr(), ws.numel(), int(torch.cuda.current_stream(dev).cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        outs.append(y.cpu().numpy())
    np.testing.assert_allclose(outs[0], outs[1], rtol=2e-6, atol=2e-6)


def test_c_host_runs_the_abi_without_torch(cuda_device):
    """tests/c/abi_run.cu: cudaMalloc + dmel_forward_f32 + dmel_backward_f32 from a plain C++ process, compared with the
    oracle's arrays stored in tests/golden/cases/*.bin (scripts/dump_case.py)."""
    import glob
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "c", "abi_run")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build_c_host()
    cases = sorted(glob.glob(os.path.join(root, "tests", "golden", "cases", "*.bin")))
    assert len(cases) >= 4
    for c in cases:
        out = subprocess.run([exe, "check", c], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and "PASS" in out.stdout, (c, out.stdout, out.stderr)


@pytest.mark.parametrize("sr,N,hop,M,windows_ms,dtype", [
    (8000, 8000, 80, 64, (10, 35, 300), torch.float32),       # the reference's grid at 8 kHz: n_fft 128 / 512 / 4096
    (16000, 16000, 160, 128, (10, 35, 300), torch.float32),   # 16 kHz: n_fft 256 / 1024 / 8192 (tiled + big-FFT paths)
    (8000, 4000, 40, 0, (35, 35.5, 20), torch.float32),       # spectrogram mode, two entries sharing one n_fft
    (8000, 8000, 80, 64, (35, 70), torch.int16),              # PCM clips, widened in the kernels
])
def test_fixed_lambda_grid_in_one_call(cuda_device, sr, N, hop, M, windows_ms, dtype):
    """SURVEY.md §8(f)-4 / search_spaces.py:22,29 (trainable=False x lambda in {10, 35, 300 ms} * sr / 6): one
    dmel_forward_multi_f32 call over the grid returns, per lambda, exactly what a separate forward returns (the per-clip
    means are summed once and shared), and the first entry matches the oracle."""
    dev = cuda_device
    g = torch.Generator().manual_seed(21)
    B = 3
    xf = torch.randn(B, N, generator=g)
    x = (xf * 3000.0).round().clamp(-32768, 32767).to(torch.int16) if dtype == torch.int16 else xf
    lams = [sr * ms * 1e-3 / 6 for ms in windows_ms]
    xd = x.to(dev)
    ys = dm.dmel_multi(xd, lams, hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=EPS)
    assert len(ys) == len(lams)
    for lam, y in zip(lams, ys):
        single = dm.dmel(xd, torch.tensor(float(lam), device=dev), hop_length=hop, optimized=True, n_mels=M,
                         sample_rate=sr, log_eps=EPS)
        assert y.shape == single.shape
        assert torch.equal(y, single), f"lambda {lam}: grid output differs from the single call"
    want = orc.port_forward(x.to(torch.float32), torch.tensor(float(lams[0])), hop=hop, optimized=True, norm=False,
                            n_mels=M, sample_rate=sr)
    check_y(ys[0].cpu().numpy(), torch.log(want + EPS).numpy(), False)


def test_layer_forward_grid(cuda_device):
    dev = cuda_device
    sr, N, hop, M = 8000, 8000, 80, 64
    x = torch.randn(2, N, generator=torch.Generator().manual_seed(3)).to(dev)
    layer = dm.MelSpectrogramLayer(torch.tensor(sr * 0.035 / 6), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop,
                                   device=str(dev), optimized=True, log_eps=EPS).to(dev)
    lams = [sr * 0.010 / 6, sr * 0.035 / 6]
    ys = layer.forward_grid(x, lams)
    with torch.no_grad():
        assert torch.equal(ys[1], layer(x))
    assert ys[0].shape == ys[1].shape == (2, 1, M, N // hop + 1)


@pytest.mark.parametrize("n_fft,M,hop,log", [(64, 32, 40, True), (32, 16, 16, True), (64, 0, 20, False), (16, 0, 3, False),
                                             (128, 64, 80, True)])
def test_tensor_core_dft_forward(cuda_device, n_fft, M, hop, log):
    """dmel_dftmm_kernels.cuh: even n_fft <= 128 run the forward as a 3xTF32 DFT-by-matrix-multiply on tcgen05 (n_fft 128
    only with DMEL_DFTMM_128=1; by default it stays on the FFT and this case checks that path).  Same oracle, same
    tolerances; several tiles of 128 frames, frames numbered across clips, a ragged last tile."""
    dev = cuda_device
    sr, N, B = 8000, 4000, 7
    lam = 0.75 * n_fft / 6
    assert orc.window_length(lam, N, True)[1] == n_fft
    g = torch.Generator().manual_seed(n_fft + M)
    x = torch.randn(B, N, generator=g) * 0.3 + 0.1
    y = dm.dmel(x.to(dev), torch.tensor(float(lam), device=dev), hop_length=hop, optimized=True, n_mels=M, sample_rate=sr,
                log_eps=EPS if log else None)
    want = orc.port_forward(x, torch.tensor(float(lam)), hop=hop, optimized=True, norm=False, n_mels=M, sample_rate=sr)
    if log:
        want = torch.log(want + EPS)
    check_y(y.cpu().numpy(), want.numpy(), not log)
    # the fused energy by-product (train.py:58) comes out of this kernel family too
    y2, energy = dm.dmel(x.to(dev), torch.tensor(float(lam), device=dev), hop_length=hop, optimized=True, n_mels=M,
                         sample_rate=sr, log_eps=EPS if log else None, return_energy=True)
    assert abs(float(energy) - float(y2.double().sum())) <= 1e-5 * float(y2.double().abs().sum())


@pytest.mark.parametrize("n_stds", [4, 8])
def test_window_support_in_standard_deviations(cuda_device, n_stds):
    """time_frequency.py:32,39: differentiable_spectrogram(..., n_stds) sizes the optimized window as
    next_power_of_2(int(n_stds * lambd)); the window formula itself does not change.  Against the reference's own
    formula evaluated with torch.stft on the CPU."""
    dev = cuda_device
    lam, N, hop = 20.0, 2000, 16
    x = torch.randn(N, generator=torch.Generator().manual_seed(n_stds))
    got, win = dm.differentiable_spectrogram(x.to(dev), torch.tensor(lam, device=dev), optimized=True, hop_length=hop,
                                             return_window=True, n_stds=n_stds)
    L = orc.next_power_of_2(np.float32(lam) * np.float32(n_stds))
    assert win.shape[0] == L
    m = torch.arange(0, L).float()
    w = torch.exp(-0.5 * torch.pow((m - L / 2) / (lam + 1e-15), 2))
    want = torch.stft(x, n_fft=L, hop_length=hop, win_length=L, window=w, return_complex=True, pad_mode='constant').abs() ** 2
    np.testing.assert_allclose(win.cpu().numpy(), w.numpy(), rtol=1e-6, atol=1e-7)
    check_y(got.cpu().numpy(), want.numpy(), True)


def test_dynamic_shared_memory_right_at_the_opt_in_threshold(cuda_device):
    """The forward-mode kernel at n_fft 1024, 64 mels, hop 80 asks for exactly 48 KB of dynamic shared memory on top of its
    64 bytes of static tables: it must opt in (the launch used to fail with 'invalid argument' — found by the n_fft sweep)."""
    dev = cuda_device
    x = torch.randn(4, 8000, generator=torch.Generator().manual_seed(1)).to(dev)
    ll = torch.tensor(128.0, device=dev, requires_grad=True)
    y = dm.dmel(x, ll, hop_length=80, optimized=True, n_mels=64, sample_rate=8000, log_eps=EPS)
    G = torch.randn(y.shape, generator=torch.Generator().manual_seed(2)).to(dev)
    (y * G).sum().backward()
    want = orc.port_forward_backward(x.cpu(), 128.0, G.cpu(), use_log=True, batched=True, hop=80, optimized=True, n_mels=64,
                                     sample_rate=8000)
    check_y(y.detach().cpu().numpy(), want[0], False)
    assert abs(float(ll.grad) - want[2]) <= G_RTOL * abs(want[2])
