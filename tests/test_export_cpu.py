"""Host-side surface that needs no GPU: export_weights_fpga / compute_layer_crc / quantize_tensor against files written by the
reference's exporter (tests/golden/make_export_fixture.py), and the drop-in modules' parameter surface."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN


def _load():
    return dict(np.load(os.path.join(GOLDEN, "ref_export.npz")))


def test_export_weights_fpga_is_byte_identical(tmp_path):
    import ofdm_gan_sr_b200.models as models
    import ofdm_gan_sr_b200.utils as utils
    r = _load()
    for tag, cls, cfg in (("g", models.MiniGenerator, None), ("d", models.MiniDiscriminator, utils.QuantizationConfig(per_channel=False))):
        m = cls()
        m.load_state_dict({k[len(tag) + 4:]: torch.as_tensor(v) for k, v in r.items() if k.startswith(tag + "_sd_")})   # interchange
        d = tmp_path / tag
        meta = utils.export_weights_fpga(m, str(d), cfg)
        ref_meta = json.loads(str(r[tag + "_metadata"]))
        assert meta == ref_meta
        assert json.loads(open(d / "metadata.json").read()) == ref_meta
        files = {k[len(tag) + 6:]: v for k, v in r.items() if k.startswith(tag + "_file_")}
        assert sorted(files) == sorted(f for f in os.listdir(d) if f.endswith(".bin"))
        for f, ref in files.items():
            assert np.array_equal(np.frombuffer(open(d / f, "rb").read(), dtype=np.uint8), ref), f
        assert utils.compute_layer_crc(next(m.parameters())) == str(r[tag + "_layer_crc"])


def test_quantize_helpers_match_reference(ref_channel):
    import ofdm_gan_sr_b200.utils as utils
    r = ref_channel
    t = torch.as_tensor(r["qt_in"])
    assert np.array_equal(utils.quantize_tensor(t, torch.tensor(1.0 / 128), 8).numpy(), r["qt_q17"])
    scale = utils.compute_scale(t, 8)
    assert float(scale) == float(r["qt_scale"])
    assert np.array_equal(utils.quantize_tensor(t, scale, 8).numpy(), r["qt_q8"])
    assert np.array_equal(utils.dequantize_tensor(utils.quantize_tensor(t, scale, 8), scale).numpy(), r["qt_deq"])
    w = torch.as_tensor(r["qt_w"])
    ws = utils.compute_scale(w, 8, per_channel=True, channel_dim=0)
    assert np.array_equal(ws.numpy(), r["qt_w_scale"]) and np.array_equal(utils.quantize_tensor(w, ws, 8).numpy(), r["qt_w_q"])
    fq = utils.FakeQuantize(8, per_channel=False).train()
    y = fq(t.clone().requires_grad_(True))
    y.sum().backward()                                        # straight-through estimator
    assert torch.allclose(y.detach(), torch.as_tensor(r["qt_deq"]), atol=1e-7)
