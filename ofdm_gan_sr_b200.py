"""Import shim: the package directory is named `ofdm-gan-sr_b200` (not a Python identifier), so
`import ofdm_gan_sr_b200` loads this file, which registers the real package under the same name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ofdm-gan-sr_b200")
_spec = importlib.util.spec_from_file_location("ofdm_gan_sr_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ofdm_gan_sr_b200"] = _mod
_spec.loader.exec_module(_mod)
