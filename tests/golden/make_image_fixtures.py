#!/usr/bin/env python3
"""Record image-path fixtures from the UNMODIFIED reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_image_fixtures.py   ->  tests/golden/ref_image.npz

ImageOFDMConverter.image_to_ofdm / ofdm_to_image (utils/ofdm_utils.py:884-994) on three synthetic images under three
converter configurations (the OFDMDataset default QPSK/8/2/16, the class default QAM16/64/16/1024, and QAM64/16/4/256), and
OFDMDataset (utils/dataset.py:38-182) over a temporary directory holding those images as PNG files: the cached clean frames,
normalisation factors and a few samples drawn after np.random.seed(3)."""
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from PIL import Image  # noqa: E402
from utils.dataset import OFDMDataset  # noqa: E402
from utils.ofdm_utils import ImageOFDMConverter  # noqa: E402

rng = np.random.default_rng(7)
images = {"gray8": rng.integers(0, 256, (8, 8), dtype=np.uint8),
          "rgb6x5": rng.integers(0, 256, (6, 5, 3), dtype=np.uint8),
          "gray100x80": rng.integers(0, 256, (80, 100), dtype=np.uint8)}     # > 4096 pixels: resized to 64x64 by the dataset
CONFIGS = {"a": ("QPSK", 8, 2, 16), "b": ("QAM16", 64, 16, 1024), "c": ("QAM64", 16, 4, 256)}
out = {"img_" + k: v for k, v in images.items()}
for tag, (mod, nsc, cp, fl) in CONFIGS.items():
    conv = ImageOFDMConverter(modulation=mod, n_subcarriers=nsc, cp_length=cp, frame_length=fl)
    for name in ("gray8", "rgb6x5"):
        iq, meta = conv.image_to_ofdm(images[name])
        rec = conv.ofdm_to_image(iq, meta["original_shape"], meta["normalization_factor"])
        out[f"{tag}_{name}_iq"] = iq
        out[f"{tag}_{name}_meta"] = np.array([meta["n_pixels"], meta["n_bits"], meta["n_qam_symbols"], meta["signal_length"]], np.int64)
        out[f"{tag}_{name}_factor"] = np.float64(meta["normalization_factor"])
        out[f"{tag}_{name}_rec"] = rec
        print(tag, name, iq.shape, meta, "reconstruction errors:", int((rec != (images[name] if images[name].ndim == 2 else
              np.dot(images[name][..., :3], [0.299, 0.587, 0.114]).astype(np.uint8))).sum()))

with tempfile.TemporaryDirectory() as d:
    for name, im in images.items():
        Image.fromarray(im).save(os.path.join(d, name + ".png"))
    ds = OFDMDataset(d, samples_per_image=4, snr_range=(5, 20))
    out["ds_files"] = np.array([p.name for p in ds.image_files])
    np.random.seed(3)
    items = [ds[i] for i in range(len(ds))]
    out["ds_len"] = np.int64(len(ds))
    out["ds_clean_cache"] = np.stack([ds._clean_signal_cache[i][0] for i in range(len(ds.image_files))])
    out["ds_factor"] = np.array([ds._clean_signal_cache[i][1]["normalization_factor"] for i in range(len(ds.image_files))])
    out["ds_noisy"] = np.stack([it["noisy"].numpy() for it in items])
    out["ds_clean"] = np.stack([it["clean"].numpy() for it in items])
    out["ds_snr"] = np.array([float(it["snr"]) for it in items])
    out["ds_loaded_gray100x80"] = ds._load_image(ds.image_files[list(out["ds_files"]).index("gray100x80.png")])
np.savez_compressed(os.path.join(HERE, "ref_image.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
