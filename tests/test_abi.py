"""The C-ABI library loads and exports every symbol include/b200grbm.h declares; without a
GPU every compute entry point fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

from image_generation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200grbm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200grbm_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.b200grbm_abi_version() == _lib.ABI_VERSION


def test_sweep_args_layout_matches_header():
    # struct_size is checked by the library at run time; here: field order / count vs the header
    text = open(os.path.join(ROOT, "include", "b200grbm.h")).read()
    body = text[text.index("typedef struct b200grbm_sweep_args"):text.index("} b200grbm_sweep_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+)(?:\[[^\]]*\])?;", body)
    assert names == [f[0] for f in _lib.SweepArgs._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_a_gpu():
    lib = _lib.load()
    a = _lib.SweepArgs()
    a.struct_size = C.sizeof(_lib.SweepArgs)
    a.n = a.n_pad = 32
    a.ell_width = 1
    a.n_tiles = 1
    a.tiles_dev = a.tile_info_dev = a.coef_dev = 16  # never dereferenced: device check comes first
    a.chains, a.chains_per_lane, a.threads, a.num_sweeps = 4, 32, 128, 1
    rc = lib.b200grbm_gibbs_sweeps(C.byref(a), None)
    assert rc != 0
    assert b"fallback" in lib.b200grbm_last_error() or b"CUDA" in lib.b200grbm_last_error()
    with pytest.raises(_lib.B200Error):
        _lib.check(rc)


def test_argument_validation_happens_before_any_launch():
    lib = _lib.load()
    a = _lib.SweepArgs()
    a.struct_size = 3
    assert lib.b200grbm_gibbs_sweeps(C.byref(a), None) == -1
    a.struct_size = C.sizeof(_lib.SweepArgs)
    a.n, a.n_pad, a.ell_width, a.n_tiles, a.chains, a.num_sweeps = 8, 8, 1, 1, 4, 1
    a.tiles_dev = a.tile_info_dev = a.coef_dev = 16
    a.chains_per_lane, a.threads = 30, 128
    assert lib.b200grbm_gibbs_sweeps(C.byref(a), None) == -2      # unsupported chains_per_lane
    a.chains_per_lane, a.threads = 32, 100
    assert lib.b200grbm_gibbs_sweeps(C.byref(a), None) == -1      # threads not a multiple of 32
    a.threads, a.chain_offset = 128, 2
    assert lib.b200grbm_gibbs_sweeps(C.byref(a), None) == -1      # chain_offset not a multiple of 4
    a.chain_offset, a.tiles_dev = 0, 8
    assert lib.b200grbm_gibbs_sweeps(C.byref(a), None) == -1      # bulk-copy source must be 16-byte aligned
    assert b"16-byte" in lib.b200grbm_last_error()


def test_sweep_smem_formula_matches_library():
    from image_generation_b200.sampler import plan_threads, sweep_smem_bytes, sweep_state_offset
    lib = _lib.load()
    for n, w, t, nt in ((5640, 15, 736, 8), (7440, 20, 480, 16), (256, 20, 128, 5), (9, 2, 64, 3), (50000, 20, 256, 200)):
        assert lib.b200grbm_sweep_smem_bytes(n, w, t, nt) == sweep_smem_bytes(n, w, t, nt)
        # the tiles' .nbr fields are byte offsets from the start of the CTA's shared memory: host and kernel must agree
        assert lib.b200grbm_sweep_state_offset(nt) == sweep_state_offset(nt)
        assert sweep_state_offset(nt) % 128 == 0 and sweep_state_offset(nt) >= 128 + 8 * nt
    # the planner never exceeds the 227 KB opt-in limit
    for n, w, sizes in ((5640, 15, [1410] * 4), (7440, 20, [1860] * 4), (40000, 20, [10000] * 4)):
        t = plan_threads(sizes, n, w)
        assert sweep_smem_bytes(n, w, t, sum(-(-s // t) for s in sizes)) <= 227 * 1024 and t % 32 == 0


def test_fp4_gram_selection_rule(monkeypatch):
    """Host-side choice of the forward Gram's operand format (image_generation_b200/mmd_tc.py): e2m1 from
    FP4_GRAM_MIN_ROWS rows up, B200GRBM_MMD_FP4 forces either form."""
    from image_generation_b200 import mmd_tc
    monkeypatch.delenv("B200GRBM_MMD_FP4", raising=False)
    assert not mmd_tc.use_fp4_gram(mmd_tc.FP4_GRAM_MIN_ROWS - 1) and mmd_tc.use_fp4_gram(mmd_tc.FP4_GRAM_MIN_ROWS)
    assert not mmd_tc.use_fp4_gram(0) and mmd_tc.use_fp4_gram()
    monkeypatch.setenv("B200GRBM_MMD_FP4", "0")
    assert not mmd_tc.use_fp4_gram(1 << 20)
    monkeypatch.setenv("B200GRBM_MMD_FP4", "1")
    assert mmd_tc.use_fp4_gram(4)
    monkeypatch.setenv("B200GRBM_MMD_FP4", "auto")          # anything else: the default rule
    assert not mmd_tc.use_fp4_gram(4) and mmd_tc.use_fp4_gram(1 << 20)


def test_sharded_mmd_exchange_mode_is_validated(monkeypatch):
    from image_generation_b200.dist import _DeviceOps
    monkeypatch.setenv("B200GRBM_MMD_EXCHANGE", "carrier-pigeon")
    with pytest.raises(ValueError):
        _DeviceOps.exchange(torch.ones(2, 4), torch.ones(2, 4), 0, 1, None)
