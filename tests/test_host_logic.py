"""CPU-side checks: host logic mirrored from the reference, filterbank packing, and that the C-ABI library
loads and exports every symbol include/dmel.h declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dmel_b200 as dm
from conftest import ROOT
from oracle import dmel_oracle as orc


def test_library_loads_and_exports_header_symbols():
    lib = dm.load()
    header = open(os.path.join(ROOT, "include", "dmel.h")).read()
    declared = set(re.findall(r"DMEL_API[^;(]*?\b(dmel_\w+)\s*\(", header))
    assert declared == set(dm._lib.EXPORTED_SYMBOLS), declared ^ set(dm._lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.dmel_abi_version() == 4
    assert lib.dmel_status_string(0) == b"ok"
    assert b"n_fft" in lib.dmel_status_string(-3)


def test_window_length_python_and_c_agree_with_reference_rule():
    lib = dm.load()
    rng = np.random.default_rng(0)
    lams = list(rng.uniform(0, 700, 300)) + [0.0, 0.16, 0.17, 1 / 3, 0.34, 42.6666, 42.67, 85.33, 85.34, 13.333, 46.667, 400.0,
                                           -46.667, 682.6, 682.7]
    for lam in lams:
        want = orc.next_power_of_2(np.float32(abs(np.float32(lam))) * np.float32(6))
        assert dm.window_length_from_lambda(lam) == want
        assert lib.dmel_window_length_from_lambda(ctypes.c_float(lam)) == want
    assert dm.next_power_of_2(0) == 2 and dm.next_power_of_2(1) == 1  # time_frequency.py:60-65 edge cases


def test_geometry_follows_reference():
    g = dm.resolve_geometry(256, 8000, 80, 46.667, True, 64)
    assert (g.win_length, g.n_fft, g.n_frames, g.n_freqs, g.rows) == (512, 512, 101, 257, 64)
    g = dm.resolve_geometry(128, 128, 1, 6.38, False, 0)        # time_frequency search space
    assert (g.win_length, g.n_fft, g.n_frames, g.rows) == (128, 256, 129, 129)
    g = dm.resolve_geometry(128, 220500, 441, 257.25, True, 128)
    assert (g.n_fft, g.n_frames) == (2048, 501)


@pytest.mark.parametrize("F,M,sr", [(257, 64, 8000), (65, 64, 8000), (1025, 128, 44100), (513, 128, 16000), (129, 16, 8000)])
def test_melbank_bit_identical_to_torchaudio_and_pack_roundtrip(F, M, sr):
    import warnings
    torchaudio = pytest.importorskip("torchaudio")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = torchaudio.functional.melscale_fbanks(n_freqs=F, f_min=0, f_max=sr // 2, n_mels=M, sample_rate=sr)
    got = dm.melscale_fbanks_htk(F, 0, sr // 2, M, sr)
    assert torch.equal(got, want)
    assert torch.equal(orc.mel_filterbank(F, 0, sr // 2, M, sr), want)
    from importlib import import_module
    mb = import_module("differentiable-mel-spectrogram_b200.melbank")
    p = mb.pack_bank(got.numpy())
    a, b = mb.unpack_bank(p)
    assert np.array_equal(a, got.numpy()) and np.array_equal(b, got.numpy())
    assert (p.bin_off[1:] - p.bin_off[:-1]).max() <= 2          # <= 2 taps per bin


def test_pack_bank_generic_matrix():
    from importlib import import_module
    mb = import_module("differentiable-mel-spectrogram_b200.melbank")
    rng = np.random.default_rng(1)
    fb = rng.random((9, 5)).astype(np.float32)
    fb[0] = 0
    fb[:, 3] = 0
    fb[4, 1] = 0
    a, b = mb.unpack_bank(mb.pack_bank(fb))
    assert np.array_equal(a, fb) and np.array_equal(b, fb)


def test_c_abi_rejects_bad_arguments_without_touching_the_device():
    lib = dm.load()
    f = lib.dmel_forward_f32
    # null x
    assert f(None, 1, 100, 100, 10.0, 64, 64, 10, 0, None, -1.0, 0, 1, 1, 1 << 20, None) == -1
    # n_fft < win_length
    assert f(8, 1, 100, 100, 10.0, 128, 64, 10, 0, None, -1.0, 0, 8, 8, 1 << 20, None) == -2
    # mel mode without a bank
    assert f(8, 1, 100, 100, 10.0, 64, 64, 10, 4, None, -1.0, 0, 8, 8, 1 << 20, None) == -1
    # workspace too small
    assert f(8, 1, 100, 100, 10.0, 64, 64, 10, 0, None, -1.0, 0, 8, 8, 16, None) == -4
    # odd n_fft has no kernel
    assert lib.dmel_nfft_supported(255) == 0 and lib.dmel_nfft_supported(512) == 1 and lib.dmel_nfft_supported(200) == 1
    assert lib.dmel_workspace_bytes(256, 512) >= 64 + 512 * 12 + 256 * 12


def test_output_affine_struct_fields_and_batchnorm_folding():
    # struct dmel_melbank: 4 int32 + 9 pointers (include/dmel.h, ABI 3: + seg_tables)
    assert ctypes.sizeof(dm._lib.MelBankStruct) == 4 * 4 + 9 * 8
    assert [f[0] for f in dm._lib.MelBankStruct._fields_][-3:] == ["out_scale", "out_shift", "seg_tables"]
    # a bank with only one of the two vectors is rejected before anything touches the device
    lib = dm.load()
    s = dm._lib.MelBankStruct()
    s.n_freqs, s.n_mels, s.mel_nnz, s.bin_nnz = 33, 4, 16, 40
    s.mel_k0 = s.mel_off = s.mel_w = s.bin_m0 = s.bin_off = s.bin_w = 16
    s.out_scale, s.out_shift = 16, None
    f = lib.dmel_forward_f32
    assert f(8, 1, 100, 100, 10.0, 64, 64, 10, 4, ctypes.byref(s), -1.0, 0, 8, 8, 1 << 20, None) == -7
    # eval-mode BatchNorm2d over mel bins == per-mel affine map (panns.py:146, 169-172)
    M = 8
    bn = torch.nn.BatchNorm2d(M).eval()
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(M, generator=g)); bn.bias.copy_(torch.randn(M, generator=g))
        bn.running_mean.copy_(torch.randn(M, generator=g)); bn.running_var.copy_(torch.rand(M, generator=g) + 0.1)
    layer = dm.MelSpectrogramLayer(torch.tensor(20.0), n_mels=M, n_points=800, sample_rate=8000, hop_length=80, optimized=True)
    layer.fold_batchnorm(bn)
    scale, shift = layer._out_affine
    y = torch.randn(3, 1, M, 11, generator=g)
    with torch.no_grad():
        want = bn(y.transpose(1, 2)).transpose(1, 2)
    assert torch.allclose(y * scale.view(1, 1, M, 1) + shift.view(1, 1, M, 1), want, rtol=1e-6, atol=1e-6)
    assert list(layer.state_dict()) == ["lambd"]
    with pytest.raises(ValueError):
        layer.fold_batchnorm(torch.nn.BatchNorm2d(M + 1))
    with pytest.raises(ValueError):
        layer.set_output_affine(torch.ones(M), torch.zeros(M - 1))
    layer.set_output_affine(None)
    assert layer._out_affine is None


def test_fft_plan_swizzle_is_xor_separable():
    """fft_device.cuh reaches the R elements of a butterfly as swz(base) ^ swz(r*stride): that needs
    swz(base + r*stride) == swz(base) ^ swz(r*stride) for every plan, pass, thread and r."""
    plans = {128: (4, [4, 4, 4, 2]), 256: (8, [8, 8, 4]), 512: (16, [8, 8, 8]), 1024: (16, [16, 8, 8]), 2048: (16, [16, 16, 8]),
             4096: (16, [16, 16, 16]), 8192: (32, [16, 16, 8, 4]), 16384: (32, [16, 16, 8, 8])}
    rot = lambda NS, R: 0 if NS >= 16 else (16 // R if NS * R >= 16 else NS)          # noqa: E731
    swz = lambda i, ro: i ^ ((ro * (i >> 4)) & 15) if ro else i                        # noqa: E731
    for L, (E, rad) in plans.items():
        GT, NS = L // E, 1
        assert int(np.prod(rad)) == L
        for p, R in enumerate(rad):
            ro_out = rot(NS, R)
            ro_in = 0 if p == 0 else rot(NS // rad[p - 1], rad[p - 1])
            STR = L // R
            seen = set()
            for tid in range(GT):
                for i in range(E // R):
                    j = tid + i * GT
                    k = j & (NS - 1)
                    j0 = (j - k) * R + k
                    for r in range(R):
                        assert swz(j0 + r * NS, ro_out) == swz(j0, ro_out) ^ swz(r * NS, ro_out)
                        assert swz(j + r * STR, ro_in) == swz(j, ro_in) ^ swz(r * STR, ro_in)
                        seen.add(swz(j0 + r * NS, ro_out))
            assert seen == set(range(L))               # the scatter is a permutation of the buffer
            NS *= R


def test_peer_and_jvp_entry_points_validate_arguments():
    lib = dm.load()
    assert lib.dmel_peer_buffer_bytes(8) == (4 * 8 + 4) * 4 and lib.dmel_peer_buffer_bytes(0) == 0 and lib.dmel_peer_buffer_bytes(33) == 0
    assert lib.dmel_peer_allreduce_f32(None, None, 0, 2, None) == -1
    assert lib.dmel_peer_allreduce_f32(8, 8, 2, 2, None) == -2          # rank outside the world
    assert lib.dmel_jvp_supported(512) == 1 and lib.dmel_jvp_supported(16384) == 0 and lib.dmel_jvp_supported(200) == 0
    assert lib.dmel_lambda_grad_f32(None, None, 10, 1.0, None, None, 0, None) == -1
    assert lib.dmel_lambda_grad_f32(8, 8, 0, 1.0, 8, 8, 64, None) == -2
    assert lib.dmel_workspace_y_sum_offset() % 8 == 0 and lib.dmel_workspace_y_sum_offset() < 64


def test_launch_plans_fit_shared_memory():
    for (B, N, L, hop, rows) in [(256, 8000, 512, 80, 64), (128, 220500, 2048, 441, 128), (1024, 160000, 1024, 160, 128),
                                 (512, 8000, 4096, 80, 64), (512, 8000, 128, 80, 64), (64, 256, 512, 1, 0), (128, 128, 256, 1, 0),
                                 (512, 8000, 256, 80, 64), (1, 50, 512, 80, 8)]:
        spec = rows == 0
        r = rows if rows else L // 2 + 1
        for which in (0, 1, 2):
            q = dm.query_launch(B, N, L, hop, r, 0 if spec else 2 * (L // 2 + 1), which)
            assert q["path"] == 0 and q["smem_bytes"] <= 227 * 1024 and q["tile_frames"] % 2 == 0, q
            assert q["threads"] in (128, 256, 512) and q["groups"] in (1, 2, 4, 8), q
    assert dm.query_launch(2, 100, 200, 5, 101, 0, 1)["path"] == 1           # non power of two -> direct DFT path
    for L, threads in ((8192, 256), (16384, 512)):                           # big-FFT path: one CTA per transform
        for which in (0, 1, 2):
            q = dm.query_launch(4, 220500, L, 441, 128, 2 * (L // 2 + 1), which)
            assert q["path"] == 2 and q["threads"] == threads and q["smem_bytes"] <= 227 * 1024, q


def test_layers_keep_the_reference_contract_and_refuse_cpu():
    layer = dm.MelSpectrogramLayer(torch.tensor(46.667), n_mels=64, n_points=8000, sample_rate=8000, hop_length=80,
                                   device="cpu", optimized=True, normalize_window=False)
    assert list(layer.state_dict().keys()) == ["lambd"] and layer.lambd.dim() == 0
    assert (layer.n_freq, layer.n_time, layer.f_max, layer.f_min) == (64, 101, 4000, 0)
    assert [n for n, _ in layer.named_parameters()] == ["lambd"]
    layer.requires_grad_(False)
    assert not layer.lambd.requires_grad
    with pytest.raises(dm.DmelError, match="no CPU"):
        layer(torch.zeros(2, 8000))
    spec = dm.SpectrogramLayer(torch.tensor(6.38), device="cpu", optimized=False, size=(129, 129), hop_length=1)
    assert list(spec.state_dict().keys()) == ["lambd"]
    with pytest.raises(dm.DmelError, match="no CPU"):
        spec(torch.zeros(2, 128))
    with pytest.raises(dm.DmelError):
        dm.dmel(torch.zeros(2, 128), torch.tensor(6.38), hop_length=1, optimized=False)


def test_lambda_mirror_follows_parameter_writes():
    layer = dm.MelSpectrogramLayer(torch.tensor(46.667), n_mels=64, n_points=8000, sample_rate=8000, hop_length=80)
    m = layer._mirror
    assert abs(m.get(layer.lambd) - 46.667) < 1e-4
    opt = torch.optim.SGD([layer.lambd], lr=1.0)
    layer.lambd.grad = torch.tensor(2.0)
    opt.step()                                           # in-place update under no_grad bumps the version counter
    assert abs(m.get(layer.lambd) - 44.667) < 1e-4
    layer.load_state_dict({"lambd": torch.tensor(90.0)})
    assert m.get(layer.lambd) == 90.0
    layer.lambd.data.fill_(13.0)                         # bypasses the version counter ...
    layer.invalidate_lambda_cache()                      # ... so the documented escape hatch is needed
    assert m.get(layer.lambd) == 13.0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "differentiable-mel-spectrogram_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn


def test_missing_library_fails_loudly_instead_of_falling_back(monkeypatch, tmp_path):
    """No CPU or PyTorch fallback exists: without libdmel_b200.so the product raises, it never computes another way."""
    lib_mod = dm._lib
    monkeypatch.setattr(lib_mod, "_lib", None)
    monkeypatch.setattr(lib_mod, "LIB_PATH", str(tmp_path / "libdmel_b200.so"))
    monkeypatch.delenv("DMEL_LIB_PATH", raising=False)
    with pytest.raises(dm.DmelError, match="no CPU fallback|not found"):
        lib_mod.load()
    layer = dm.MelSpectrogramLayer(torch.tensor(20.0), n_mels=8, n_points=800, sample_rate=8000, hop_length=80, optimized=True)
    with pytest.raises(dm.DmelError):
        layer(torch.randn(2, 800))              # CPU tensors are refused as well


def test_folded_dft_with_3xtf32_split_is_fp32_accurate():
    """The formulation of csrc/dmel_dftmm_kernels.cuh, restated in numpy: a frame folded around its centre
    (e[j] = xw[j] + xw[L-j], o[j] = xw[j] - xw[L-j]) needs two real GEMMs with K = L/2, and operands split x = hi + lo
    with hi = tf32(x) give fp32 accuracy from three tf32 products (the tensor core reads the upper 19 bits of lo)."""
    def tf32_rna(a):                      # round to nearest, 10-bit mantissa (cvt.rna.tf32.f32)
        u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
        u = (u + 0x1000) & 0xFFFFE000
        return u.astype(np.uint32).view(np.float32)

    def tf32_trunc(a):                    # what the MMA does to an fp32 operand
        return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    rng = np.random.default_rng(3)
    for L in (128, 64, 100, 36):
        H = L // 2
        xw = (rng.standard_normal((40, L)) * np.exp(-0.5 * ((np.arange(L) - L / 2) / (L / 6)) ** 2)).astype(np.float32)
        j = np.arange(H)
        e = xw[:, :H].copy()
        e[:, 1:] += xw[:, :H:-1][:, : H - 1]              # xw[L-j], j = 1..H-1
        o = xw[:, :H].copy()
        o[:, 1:] -= xw[:, :H:-1][:, : H - 1]
        k = np.arange(H + 1)
        ang = 2 * np.pi * np.outer(k, j) / L
        cosm, sinm = np.cos(ang).astype(np.float32), (-np.sin(ang)).astype(np.float32)

        def mm3(a, b):                    # A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation
            ah, bh = tf32_rna(a), tf32_rna(b)
            al, bl = tf32_trunc(a - ah), tf32_trunc(b - bh)
            return (ah @ bh.T + al @ bh.T + ah @ bl.T).astype(np.float32)

        re = mm3(e, cosm) + np.outer(xw[:, H], (-1.0) ** k).astype(np.float32)
        im = mm3(o, sinm)
        want = np.fft.rfft(xw.astype(np.float64), axis=1)
        scale = np.abs(want).max()
        assert np.abs(re - want.real).max() <= 2e-6 * scale, L
        assert np.abs(im - want.imag).max() <= 2e-6 * scale, L
        # a single tf32 pass (no split) is three orders of magnitude worse: the reason for the split
        one = (tf32_rna(e) @ tf32_rna(cosm).T).astype(np.float32) + np.outer(xw[:, H], (-1.0) ** k).astype(np.float32)
        assert np.abs(one - want.real).max() > 1e-4 * scale
