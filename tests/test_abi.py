"""The C-ABI library loads on a CPU-only machine and exports every symbol include/ofdmgan.h declares.
No compute call is made here (there is no GPU): only symbol / layout / error-path checks."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "ofdmgan.h")


@pytest.fixture(scope="module")
def pkg():
    import ofdm_gan_sr_b200 as p
    if not os.path.exists(p._lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return p


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ofdmgan_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    names = _declared()
    for n in ("ofdmgan_gen_fwd_f32", "ofdmgan_gen_bwd_f32", "ofdmgan_gen_fwd_q", "ofdmgan_chan_sim", "ofdmgan_sim_gen_metrics",
              "ofdmgan_sim_gen_metrics_host", "ofdmgan_critic_step", "ofdmgan_gen_step", "ofdmgan_gen_step_fake", "ofdmgan_gen_train_ctr", "ofdmgan_adam",
              "ofdmgan_gradient_penalty", "ofdmgan_disc_fwd_f32", "ofdmgan_disc_bwd_f32"):
        assert n in names


def test_library_exports_every_declared_symbol(pkg):
    L = ctypes.CDLL(pkg._lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), f"libofdmgan.so does not export {name}"
    assert set(pkg._lib.EXPORTS) == set(_declared())       # the ctypes table covers the whole header


def test_abi_version_and_error_strings(pkg):
    L = pkg._lib.lib()
    assert L.ofdmgan_abi_version() == 16
    assert L.ofdmgan_error_string(0) == b"ok"
    assert b"invalid argument" in L.ofdmgan_error_string(-1)
    assert b"streams" in L.ofdmgan_error_string(-2)
    assert b"not supported" in L.ofdmgan_error_string(-3)


def test_struct_layout_matches_header(pkg):
    # ofdmgan_chan_cfg: 14 x 4-byte fields + (4 x 4) + int64 + 2 x int32, natural alignment
    assert ctypes.sizeof(pkg._lib.ChanCfg) == 176
    assert pkg._lib.ChanCfg.frames_per_snr.offset == 80
    assert ctypes.sizeof(pkg._lib.ChanRand) == 8 * ctypes.sizeof(ctypes.c_void_p)
    import oracle
    assert ctypes.sizeof(oracle.ChanCfg) == 176
    for (n1, t1), (n2, t2) in zip(pkg._lib.ChanCfg._fields_, oracle.ChanCfg._fields_):
        assert n1 == n2 and ctypes.sizeof(t1) == ctypes.sizeof(t2)


def test_argument_errors_need_no_device(pkg):
    """NULL pointers / bad sizes are rejected before any CUDA call."""
    L = pkg._lib.lib()
    assert L.ofdmgan_gen_fwd_f32(None, None, None, 4, 0.2, None) == -1
    assert L.ofdmgan_critic_step(None, None, None, None, 0, 0, 0, None, 10.0, 0.2, 4, 4, None, None) == -1
    assert L.ofdmgan_adam(None, None, None, None, 10, 2e-4, 0.0, 0.9, 1e-8, 1, 1.0, None) == -1
    assert L.ofdmgan_gen_step_fake(None, None, None, None, None, 1.0, 100.0, 0.2, 4, 4, None, None) == -1
    assert L.ofdmgan_gen_train_ctr(None, None, None, None, None, None, None, 2e-4, 0.0, 0.9, 1e-8, None, 1.0, 100.0, 0.2, 4, 4, None, 0, 0,
                                   None, None) == -1
    assert L.ofdmgan_critic_train_ctr(None, None, None, 0, 0, None, None, None, None, 2e-4, 0.0, 0.9, 1e-8, 10.0, 0.2, 4, 4, None, 0, None,
                                      None) == -1
    cfg = pkg.ops.make_cfg(normalize=7)
    assert L.ofdmgan_chan_sim(ctypes.byref(cfg), None, 0, 0, None, None, None, 4, None) == -1
    cfg = pkg.ops.make_cfg(n_fft=64)
    assert L.ofdmgan_chan_sim(ctypes.byref(cfg), None, 0, 0, None, None, None, 4, None) == -3


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(pkg.OfdmGanError):
        pkg.ops.gen_fwd_f32(torch.zeros(4, 2, 16), torch.zeros(258))
    with pytest.raises((pkg.OfdmGanError, RuntimeError, AssertionError)):
        pkg.ops.chan_sim(pkg.ops.make_cfg(), 4)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "ofdm-gan-sr_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


def test_flat_parameter_cache_belongs_to_the_tensor_objects(pkg):
    """ops.flat_cached keys on (address, version); a NEW tensor that reuses a freed tensor's address with the same version counter
    (torch's caching allocator does that between two freshly built modules) must not get the old tensor's values."""
    import torch
    ops = pkg.ops
    ops._FLAT_CACHE.clear()
    a = [torch.arange(6.0), torch.ones(3)]
    fa = ops.flat_cached(a)
    assert ops.flat_cached(a) is fa                                       # unchanged parameters: the cached vector
    a[0].mul_(2.0)
    assert torch.equal(ops.flat_cached(a), torch.cat([x.reshape(-1) for x in a]))     # in-place write: rebuilt
    # the same storage and version counters seen through different tensor objects = what address reuse looks like to the key
    key = tuple((p.data_ptr(), p._version) for p in a)
    b = [torch.empty(0).set_(p.untyped_storage(), 0, p.shape) for p in a]
    stale = torch.full((9,), -1.0)
    import weakref
    dead = [torch.zeros(1), torch.zeros(1)]
    ops._FLAT_CACHE[tuple(k[0] for k in key)] = (tuple((p.data_ptr(), p._version) for p in b), stale, tuple(weakref.ref(d) for d in dead))
    del dead
    got = ops.flat_cached(b)
    assert got is not stale and torch.equal(got, torch.cat([x.reshape(-1) for x in b]))
