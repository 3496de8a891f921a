"""CPU suite: the product package's HOST LOGIC (kernel sequencing, flags, samplers, quirks) run
against the golden fixtures with the C ABI emulated on the CPU (tests/emu_lib.py, test-only)."""
import ctypes
import os
import re

import pytest
import torch

import emu_lib
import parity_cases as C


@pytest.fixture
def emu(monkeypatch):
    return emu_lib.install(monkeypatch)


def test_setup_bit_exact():
    C.case_setup_bit_exact()


def test_fft(emu):
    C.case_fft("cpu")


def test_sense_and_prox(emu):
    C.case_prox("cpu")
    C.case_tv("cpu")


def test_scorenet_small(emu):
    C.case_scorenet_small("cpu")


def test_operand_shift_host_logic(emu):
    """the plan's scale bookkeeping (which operand carries which power of two) and the self-escalating first forward"""
    C.case_operand_shift("cpu")


def test_pooled_convs_in_their_strided_forms(emu, monkeypatch):
    """ConvMeanPool as one 4x4 stride-2 convolution on space-to-depth operands (weights combined on the host, 16 of 36 blocks
    non-zero) and the pooled 1x1 shortcut on pooled operands: the plan takes these paths and gives the score of the
    conv-then-pool order (IPDM_POOL_AFTER_CONV / IPDM_POOL_AFTER_SHORTCUT) up to operand rounding."""
    cfg = C.make_config("ACDC", 8, 32, 12, 30.0)
    x, y = (C.rrand(1301, 2, 1, 32, 32) * 3 - 1), torch.tensor([0, 7])
    net0, _ = C.build_net(C.NCSNv2Deepest, "NCSNv2Deepest_ngf8", 1, cfg, "cpu")
    net0(x, y)                      # a problem this small keeps the pool-after-conv order (less than one wave of work items)
    assert not any(k.endswith(".xp16") for k in next(iter(net0._plans.values())).bufs)
    monkeypatch.setenv("IPDM_POOL_STRIDED_ALWAYS", "1")
    net, _ = C.build_net(C.NCSNv2Deepest, "NCSNv2Deepest_ngf8", 1, cfg, "cpu")
    out = net(x, y)
    plan = next(iter(net._plans.values()))
    assert any(k.endswith(".conv2.conv.s2d") for k in plan.w) and any(k.endswith(".xp16") for k in plan.bufs)
    monkeypatch.delenv("IPDM_POOL_STRIDED_ALWAYS")
    monkeypatch.setenv("IPDM_POOL_AFTER_CONV", "1")
    monkeypatch.setenv("IPDM_POOL_AFTER_SHORTCUT", "1")
    net2, _ = C.build_net(C.NCSNv2Deepest, "NCSNv2Deepest_ngf8", 1, cfg, "cpu")
    ref = net2(x, y)
    assert not any(k.endswith(".xp16") for k in next(iter(net2._plans.values())).bufs)
    assert C.rel_l2(out, ref) < 1e-3, C.rel_l2(out, ref)


def test_ncsn3d_shallow_host_logic(emu):
    """kernel sequence of the 3-D temporal prior (slice-shifted conv launches, gathers, pools) against the reference output"""
    C.case_ncsn3d_shallow("cpu")


def test_metrics_and_result_files(emu):
    C.case_metrics("cpu")


def test_sampler_uncond(emu):
    C.case_sampler_uncond("cpu")


def test_sampler_sense(emu):
    C.case_sampler_sense("cpu")


def test_sampler_cine(emu):
    C.case_sampler_cine("cpu")


def test_full_chain_posterior_metrics(emu):
    C.case_full_chain_metrics("cpu", levels=12)


def test_map_baselines(emu):
    C.case_map_baselines("cpu")


def test_deeper_langevin_seg(emu):
    C.case_deeper_langevin_seg("cpu")


def test_no_cpu_fallback():
    """Without the emulator the product refuses CPU tensors instead of silently computing on the host."""
    from inverseproblemwithdiffusionmodel_b200 import _lib
    from inverseproblemwithdiffusionmodel_b200.ncsn.linear_transforms import i2k_complex
    with pytest.raises(_lib.IpdmError):
        i2k_complex(torch.zeros(1, 1, 8, 8, dtype=torch.complex64))


def test_library_exports_every_declared_symbol():
    """libipdm_b200.so loads and exports every function include/ipdm_b200.h declares (no compute calls)."""
    from inverseproblemwithdiffusionmodel_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ipdm_b200.h")).read()
    declared = set(re.findall(r"\b(ipdm_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ipdm_b200.h")).read()
    declared_version = int(re.search(r"ipdm_abi_version\(void\);\s*/\*\s*(\d+)\s*\*/", header).group(1))
    assert _lib.lib().ipdm_abi_version() == _lib.ABI_VERSION == declared_version == 3     # header, binding and library agree
    assert _lib.lib().ipdm_sense_workspace_bytes(4, 2, 256, 256) == 4 * 2 * 256 * 256 * 8


def test_graft_entry_build_check_agrees_with_the_binding():
    """__graft_entry__.build() ends with an ABI check: it must compare against the binding's version, not a literal."""
    import inspect
    import __graft_entry__
    src = inspect.getsource(__graft_entry__.build)
    assert "_lib.ABI_VERSION" in src


def test_noise_uniform_conversion_stays_inside_the_open_interval():
    """The in-kernel Box-Muller draws take u = (k + 0.5) * 2^-23 from 23 random bits (csrc/common.cuh:u01_open).  In fp32
    that is exact for every k, so 0 < u < 1 and log(u) < 0 always.  The first version used 24 bits: k + 0.5 is not
    representable above 2^23, the top value rounds to 2^24 and u = 1.0 -- radius 0 * rsqrt(0) = NaN about once per 1.7e7
    draws, which no short parity case can see (found by a full 6933-step chain).  This pins the arithmetic fact."""
    import numpy as np
    k = np.arange(1 << 23, dtype=np.uint32)
    u = (k.astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)
    assert u.dtype == np.float32
    assert float(u.min()) == 2.0 ** -24 and float(u.max()) == 1.0 - 2.0 ** -24
    assert np.array_equal(u.astype(np.float64), (k.astype(np.float64) + 0.5) / 8388608.0)      # exact, no rounding anywhere
    top24 = (np.float32((1 << 24) - 1) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)      # what 24 bits would give
    assert float(top24) == 1.0
