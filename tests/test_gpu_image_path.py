"""GPU parity, image path: ImageOFDMConverter / OFDMDataset mirrors (bits -> QAM -> OFDM kernels, channel kernel with the
transmit frame injected) vs fixtures recorded from the reference (tests/golden/make_image_fixtures.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, assert_close

pytestmark = pytest.mark.gpu

CONFIGS = {"a": ("QPSK", 8, 2, 16), "b": ("QAM16", 64, 16, 1024), "c": ("QAM64", 16, 4, 256)}


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLDEN, "ref_image.npz"))


@pytest.fixture(scope="module")
def utils():
    import ofdm_gan_sr_b200.utils as u
    assert torch.cuda.is_available()
    return u


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("name", ["gray8", "rgb6x5"])
def test_image_to_ofdm_and_back(utils, ref, tag, name):
    mod, nsc, cp, fl = CONFIGS[tag]
    conv = utils.ImageOFDMConverter(modulation=mod, n_subcarriers=nsc, cp_length=cp, frame_length=fl)
    iq, meta = conv.image_to_ofdm(ref["img_" + name])
    assert iq.dtype == np.float32 and iq.shape == (2, fl)
    assert_close(iq, ref[f"{tag}_{name}_iq"], 2e-6, "image_to_ofdm")
    assert [meta["n_pixels"], meta["n_bits"], meta["n_qam_symbols"], meta["signal_length"]] == list(ref[f"{tag}_{name}_meta"])
    assert abs(meta["normalization_factor"] - float(ref[f"{tag}_{name}_factor"])) <= 2e-6 * float(ref[f"{tag}_{name}_factor"])
    assert tuple(meta["original_shape"]) == ref[f"{tag}_{name}_rec"].shape
    # decode the REFERENCE's signal: same pixels as the reference's own ofdm_to_image (incl. its truncation losses in config a)
    # Pixels whose 8 bits all come from transmitted symbols within the kept frame are determined; the rest are decisions on
    # all-zero padding carriers, i.e. on the rounding noise of fft(ifft(.)) (1e-16 in the reference, 1e-7 here): not compared.
    kept_syms = min(meta["n_qam_symbols"], (fl // conv.ofdm.samples_per_symbol) * conv.ofdm.n_data_subcarriers)
    det = kept_syms * conv.qam.bits_per_symbol // 8
    want = ref[f"{tag}_{name}_rec"].reshape(-1)
    rec = conv.ofdm_to_image(ref[f"{tag}_{name}_iq"], meta["original_shape"], float(ref[f"{tag}_{name}_factor"]))
    assert rec.dtype == np.uint8 and rec.shape == ref[f"{tag}_{name}_rec"].shape and det > 0
    assert np.array_equal(rec.reshape(-1)[:det], want[:det])
    # and our own signal decodes to the same pixels
    own = conv.ofdm_to_image(iq, meta["original_shape"], meta["normalization_factor"])
    assert np.array_equal(own.reshape(-1)[:det], want[:det])
    if tag != "a":                                             # nothing truncated: the image itself comes back
        gray = conv._gray(ref["img_" + name]).reshape(-1)
        assert np.array_equal(own.reshape(-1)[:det], gray[:det]) and det >= gray.size - 1


def test_batched_conversion_equals_single(utils, ref):
    conv = utils.ImageOFDMConverter(modulation="QPSK", n_subcarriers=8, cp_length=2, frame_length=16)
    ims = [ref["img_gray8"], ref["img_rgb6x5"]]
    iq, factor, metas = conv.images_to_ofdm(ims)
    for i, im in enumerate(ims):
        one, meta = conv.image_to_ofdm(im)
        assert np.array_equal(iq[i].cpu().numpy(), one) and abs(float(factor[i]) - meta["normalization_factor"]) < 1e-6
        assert metas[i]["n_bits"] == meta["n_bits"]


def _write_images(ref, d):
    from PIL import Image
    for name in ("gray8", "rgb6x5", "gray100x80"):
        Image.fromarray(ref["img_" + name]).save(os.path.join(d, name + ".png"))


def test_ofdm_dataset_matches_reference(utils, ref, tmp_path):
    _write_images(ref, str(tmp_path))
    ds = utils.OFDMDataset(str(tmp_path), samples_per_image=4, snr_range=(5, 20), seed=1)
    assert [p.name for p in ds.image_files] == list(ref["ds_files"]) and len(ds) == int(ref["ds_len"]) == 12
    assert np.array_equal(ds._load_image(ds.image_files[0]), ref["ds_loaded_gray100x80"])       # PIL LANCZOS resize to 64x64
    clean, factor = ds.clean_frames()
    assert_close(clean.cpu().numpy(), ref["ds_clean_cache"], 2e-6, "cached clean frames")
    assert_close(factor.cpu().numpy(), ref["ds_factor"], 2e-6, "normalisation factors")
    b = ds.batch(0, 12)
    assert b["noisy"].shape == (12, 2, 16) and b["clean"].shape == (12, 2, 16) and b["snr"].shape == (12,)
    snr = b["snr"].cpu().numpy()
    assert (snr >= 5).all() and (snr <= 20).all() and len(np.unique(snr)) == 12
    n, c = b["noisy"].cpu().numpy(), b["clean"].cpu().numpy()
    # joint normalisation of dataset.py:143-147: max(max|noisy|, max|clean|) == 1, clean = cached clean / that maximum
    m = np.maximum(np.abs(n).max(axis=(1, 2)), np.abs(c).max(axis=(1, 2)))
    assert np.allclose(m, 1.0, atol=1e-6)
    for i in range(12):
        k = c[i].ravel() @ ref["ds_clean_cache"][i // 4].ravel() / (ref["ds_clean_cache"][i // 4].ravel() ** 2).sum()
        assert 0 < k <= 1 + 1e-6 and np.allclose(c[i], k * ref["ds_clean_cache"][i // 4], atol=1e-6)
        # the same relation holds in the reference's samples
        kr = ref["ds_clean"][i].ravel() @ ref["ds_clean_cache"][i // 4].ravel() / (ref["ds_clean_cache"][i // 4].ravel() ** 2).sum()
        assert np.allclose(ref["ds_clean"][i], kr * ref["ds_clean_cache"][i // 4], atol=1e-6)
    item = ds[5]
    assert torch.equal(item["noisy"], b["noisy"][5]) and torch.equal(item["clean"], b["clean"][5])


def test_ofdm_dataset_noise_level(utils, ref, tmp_path):
    """The residual noisy - factor * clean has the AWGN power the drawn SNR implies (measured per-frame signal power)."""
    _write_images(ref, str(tmp_path))
    ds = utils.OFDMDataset(str(tmp_path), samples_per_image=4096, snr_range=(10, 10), seed=2)
    clean_all, factor_all = ds.clean_frames()
    b = ds.batch(0, len(ds))
    n, c = b["noisy"].double(), b["clean"].double()
    idx = torch.arange(len(ds), device=n.device) // 4096
    fac = factor_all[idx].double()[:, None, None]
    resid = n - c * fac                                        # both were divided by the same maximum
    p_sig = ((c * fac) ** 2).sum(dim=(1, 2)) / 16
    p_noise = (resid ** 2).sum(dim=(1, 2)) / 16
    for im in range(3):
        sel = idx == im
        snr_est = 10 * torch.log10(p_sig[sel].mean() / p_noise[sel].mean())
        assert abs(float(snr_est) - 10.0) < 0.15, (im, float(snr_est))
    # the reference's samples obey the same relation at their own SNRs (sanity of the fixture)
    r = ref["ds_noisy"].astype(np.float64) - ref["ds_clean"].astype(np.float64) * ref["ds_factor"][np.arange(12) // 4][:, None, None]
    assert np.isfinite(r).all()


def test_loader_over_image_dataset(utils, ref, tmp_path):
    _write_images(ref, str(tmp_path))
    ds = utils.OFDMDataset(str(tmp_path), samples_per_image=8)
    loader = utils.create_dataloader(ds, batch_size=6)
    batches = list(loader)
    assert len(batches) == 4 and all(bt["noisy"].shape == (6, 2, 16) and bt["noisy"].is_cuda for bt in batches)
    with pytest.raises(IndexError):
        ds[len(ds)]
