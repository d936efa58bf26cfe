#!/usr/bin/env python3
"""Record the reference's export_weights_fpga output for a seeded MiniGenerator / MiniDiscriminator (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_export_fixture.py   ->  tests/golden/ref_export.npz

Stores the state_dicts, the metadata.json the reference wrote (as a JSON string) and every .bin file's bytes."""
import json
import os
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import torch  # noqa: E402
from models import MiniDiscriminator, MiniGenerator  # noqa: E402
from utils.quantization import QuantizationConfig, compute_layer_crc, export_weights_fpga  # noqa: E402

out = {}
for tag, cls, cfg in (("g", MiniGenerator, None), ("d", MiniDiscriminator, QuantizationConfig(per_channel=False))):
    torch.manual_seed(3)
    m = cls()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.uniform_(-0.3, 0.3)
    for k, v in m.state_dict().items():
        out[f"{tag}_sd_{k}"] = v.numpy().copy()
    with tempfile.TemporaryDirectory() as d:
        export_weights_fpga(m, d, cfg)
        out[f"{tag}_metadata"] = np.array(open(os.path.join(d, "metadata.json")).read())
        for f in sorted(os.listdir(d)):
            if f.endswith(".bin"):
                out[f"{tag}_file_{f}"] = np.frombuffer(open(os.path.join(d, f), "rb").read(), dtype=np.uint8)
    out[f"{tag}_layer_crc"] = np.array(compute_layer_crc(next(m.parameters())))
np.savez_compressed(os.path.join(HERE, "ref_export.npz"), **out)
print(sorted(out)[:12], len(out))
