"""SyntheticOFDMDataset with the reference's interface (utils/dataset.py:185-293), generated on the GPU by kernel (1).

The reference builds one sample per `__getitem__` call in NumPy from the process-global np.random state (about 90-160 us
per frame, the dominant cost of train.py --synthetic).  Here a sample is a pure function of (seed, frame index) through
Philox4x32-10 counters, so any index range can be produced in one launch, on any rank, in any order.
"""
from typing import Dict, Iterator, List, Tuple

import torch

from .. import ops
from .._lib import OfdmGanError


class SyntheticOFDMDataset(torch.utils.data.Dataset):
    """Same constructor as utils/dataset.py:195-206, plus `seed` and `device`.

    __getitem__(idx) -> {'noisy': [2,16] f32, 'clean': [2,16] f32, 'snr': scalar f32} (CUDA tensors)
    batch(start, B)  -> the same for frames start..start+B-1 in one launch ([B,2,16], [B,2,16], [B])
    """

    def __init__(self, n_samples: int = 10000, frame_length: int = 16, snr_range: Tuple[float, float] = (0, 30),
                 channel_type: str = "awgn", nonlinear: bool = False, pa_saturation: float = 1.0, iq_imbalance_db: float = 1.0,
                 iq_phase_deg: float = 5.0, phase_noise_dbchz: float = -80, seed: int = 0, device=None, symbol_source: str = "gaussian"):
        if frame_length != 16:
            raise OfdmGanError("libofdmgan builds 16-sample frames only (the MiniGenerator / RTL frame length)")
        if channel_type.lower() not in ops.CHANNEL_TYPES:
            raise ValueError(f"Unknown channel type: {channel_type}")
        self.n_samples, self.frame_length, self.snr_range = n_samples, frame_length, tuple(snr_range)
        self.nonlinear, self.pa_saturation = nonlinear, pa_saturation
        self.iq_imbalance_db, self.iq_phase_deg, self.phase_noise_dbchz = iq_imbalance_db, iq_phase_deg, phase_noise_dbchz
        self.seed, self.device, self.epoch = seed, device, 0
        src = dict(gaussian=dict(), qpsk=dict(symbol_source=ops.SYM_QPSK, n_fft=16, cp_len=0, ifft_scale=ops.SCALE_SQRT_N))[symbol_source]
        self.cfg = ops.make_cfg(nonlinear=nonlinear, pa_saturation=pa_saturation, iq_imbalance_db=iq_imbalance_db,
                                iq_phase_deg=iq_phase_deg, phase_noise_dbchz=phase_noise_dbchz, snr_mode=ops.SNR_UNIFORM,
                                snr_lo=float(snr_range[0]), snr_hi=float(snr_range[1]), normalize=ops.NORM_JOINT,
                                channel_type=channel_type.lower(), **src)

    def __len__(self) -> int:
        return self.n_samples

    def set_epoch(self, epoch: int):
        """Fresh frames every epoch, like the reference (which draws new randomness on every __getitem__)."""
        self.epoch = epoch

    def batch(self, start: int, B: int) -> Dict[str, torch.Tensor]:
        clean, noisy, snr = ops.chan_sim(self.cfg, B, seed=self.seed, frame0=self.epoch * self.n_samples + start, device=self.device)
        return {"noisy": noisy, "clean": clean, "snr": snr}

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        if idx < 0 or idx >= self.n_samples:
            raise IndexError(idx)
        b = self.batch(idx, 1)
        return {"noisy": b["noisy"][0], "clean": b["clean"][0], "snr": b["snr"][0]}


class OFDMDataset(torch.utils.data.Dataset):
    """utils/dataset.py:38-182: (noisy, clean) pairs made from image files.  Same constructor, plus `seed` / `device`.

    The clean frame of an image (first `frame_length` samples of its OFDM signal, normalised to [-1, 1]) is computed once
    for the whole directory on the GPU; a sample then is that frame, rescaled, through the channel at a uniform SNR
    (kernel (1) with the transmit frame injected), and the reference's joint normalisation
    max(max|noisy|, max|clean_normalised|) (dataset.py:143-147 - the clean frame is *not* rescaled there; kept).
    Decoding the files stays with PIL, exactly as in the reference (`_load_image`, dataset.py:168-182)."""

    def __init__(self, image_dir: str, frame_length: int = 16, modulation: str = "QPSK", n_subcarriers: int = 8, cp_length: int = 2,
                 snr_range: Tuple[float, float] = (0, 30), channel_type: str = "awgn", samples_per_image: int = 10, transform=None,
                 seed: int = 0, device=None):
        from pathlib import Path

        from .ofdm_utils import ImageOFDMConverter
        if frame_length != 16:
            raise OfdmGanError("libofdmgan builds 16-sample frames only (the MiniGenerator / RTL frame length)")
        if channel_type.lower() not in ops.CHANNEL_TYPES:
            raise ValueError(f"Unknown channel type: {channel_type}")
        self.image_dir, self.frame_length, self.snr_range = Path(image_dir), frame_length, tuple(snr_range)
        self.channel_type, self.samples_per_image, self.transform = channel_type.lower(), samples_per_image, transform
        self.seed, self.device, self.epoch = seed, device, 0
        self.converter = ImageOFDMConverter(modulation=modulation, n_subcarriers=n_subcarriers, cp_length=cp_length, frame_length=frame_length)
        self.image_files = self._find_images()
        self.cfg = ops.make_cfg(pa=False, iq=False, pn=False, snr_mode=ops.SNR_UNIFORM, snr_lo=float(snr_range[0]),
                                snr_hi=float(snr_range[1]), normalize=ops.NORM_NONE, channel_type=self.channel_type)
        self._cfg_joint = ops.make_cfg(pa=False, iq=False, pn=False, snr_mode=ops.SNR_UNIFORM, snr_lo=float(snr_range[0]),
                                       snr_hi=float(snr_range[1]), normalize=ops.NORM_JOINT, channel_type=self.channel_type)
        self._clean, self._factor = None, None

    def _find_images(self):
        images = []
        if self.image_dir.exists():
            for ext in (".png", ".jpg", ".jpeg", ".bmp", ".tiff"):
                images.extend(self.image_dir.glob(f"*{ext}"))
                images.extend(self.image_dir.glob(f"*{ext.upper()}"))
        return sorted(images)

    def _load_image(self, path):
        import numpy as np
        from PIL import Image
        image = Image.open(path)
        if image.mode != "L":
            image = image.convert("L")
        if image.size[0] * image.size[1] > 4096:
            image = image.resize((64, 64), Image.Resampling.LANCZOS)
        return np.array(image)

    def clean_frames(self):
        """([n_images, 2, 16] normalised clean frames, [n_images] normalisation factors), device-resident, built once"""
        if self._clean is None:
            self._clean, self._factor, _ = self.converter.images_to_ofdm([self._load_image(p) for p in self.image_files])
        return self._clean, self._factor

    def __len__(self) -> int:
        return len(self.image_files) * self.samples_per_image

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def batch(self, start: int, B: int) -> Dict[str, torch.Tensor]:
        clean_all, factor_all = self.clean_frames()
        idx = torch.arange(start, start + B, device=clean_all.device) // self.samples_per_image
        clean_n, factor = clean_all[idx], factor_all[idx]
        # one launch: the cached frame goes through the channel at signal scale (tx * tx_gain), the clean output stays the cached
        # frame, and both are divided by max(max|noisy|, max|clean|) inside the kernel (dataset.py:131-147)
        cfg = self._cfg_joint
        clean, noisy, snr = ops.chan_sim(cfg, B, seed=self.seed, frame0=self.epoch * len(self) + start, device=clean_all.device,
                                         tx=clean_n.reshape(B, 32).contiguous(), tx_gain=factor.contiguous())
        if self.transform:
            noisy, clean = self.transform(noisy, clean)
        return {"noisy": noisy, "clean": clean, "snr": snr}

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        if idx < 0 or idx >= len(self):
            raise IndexError(idx)
        b = self.batch(idx, 1)
        return {"noisy": b["noisy"][0], "clean": b["clean"][0], "snr": b["snr"][0]}


class GPUBatchLoader:
    """Iterates a SyntheticOFDMDataset in device-resident batches: the replacement for DataLoader(num_workers=0) at
    train.py:660 (the dict it yields is what train.py:327-329 consumes; `.to(device)` is then a no-op).

    rank / world_size shard each global batch by contiguous frame ranges (no exchange needed).  Every rank always gets the SAME
    local batch size - the data-parallel step scales gradients by 1 / (B_local x world): with drop_last=False the ragged last global
    batch is split evenly and its up to world-1 leftover frames are dropped."""

    def __init__(self, dataset: SyntheticOFDMDataset, batch_size: int = 32, drop_last: bool = True, rank: int = 0, world_size: int = 1):
        self.dataset, self.batch_size, self.drop_last, self.rank, self.world = dataset, batch_size, drop_last, rank, world_size
        self._epoch = 0

    def __len__(self) -> int:
        n, g = len(self.dataset), self.batch_size * self.world
        tail = 0 if self.drop_last else (1 if (n - (n // g) * g) // self.world > 0 else 0)
        return n // g + tail

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        self.dataset.set_epoch(self._epoch)
        self._epoch += 1
        n, g = len(self.dataset), self.batch_size * self.world
        full = n // g
        for i in range(full):
            yield self.dataset.batch(i * g + self.rank * self.batch_size, self.batch_size)
        per = (n - full * g) // self.world                          # the ragged last global batch, in equal shards
        if not self.drop_last and per > 0:
            yield self.dataset.batch(full * g + self.rank * per, per)


def create_dataloader(dataset, batch_size: int = 32, shuffle: bool = True, num_workers: int = 4, drop_last: bool = True,
                      rank: int = 0, world_size: int = 1):
    """Signature of utils/dataset.py:296-323.  Samples are i.i.d. functions of their index, so `shuffle` changes nothing
    statistically and `num_workers` is unused: batches are produced by one kernel launch each."""
    return GPUBatchLoader(dataset, batch_size=batch_size, drop_last=drop_last, rank=rank, world_size=world_size)


def generate_test_samples(n_samples: int = 100, frame_length: int = 16, snr_values: List[float] = (5, 10, 15, 20, 25),
                          channel_type: str = "awgn", seed: int = 0, device=None) -> Dict[float, List[Dict[str, torch.Tensor]]]:
    """Signature and result structure of utils/dataset.py:326-383 (SNR -> list of {'noisy','clean','snr'}): Gaussian-symbol
    frames through AWGN at fixed SNRs with joint normalisation, one launch per SNR value; the per-sample dicts are views
    into the batched device tensors."""
    if frame_length != 16 or channel_type != "awgn":
        raise OfdmGanError("generate_test_samples: only frame_length 16 / 'awgn' are built")
    out = {}
    for i, snr in enumerate(snr_values):
        cfg = ops.make_cfg(snr_mode=ops.SNR_GRID, snr_lo=float(snr), snr_step=0.0, n_snr=1, normalize=ops.NORM_JOINT)
        clean, noisy, _ = ops.chan_sim(cfg, n_samples, seed=seed, frame0=i * n_samples, device=device, want_snr=False)
        out[snr] = [{"noisy": noisy[j], "clean": clean[j], "snr": snr} for j in range(n_samples)]
    return out
