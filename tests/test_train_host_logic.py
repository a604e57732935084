"""CPU: the host-side plumbing of the training step that needs no device — the BatchNorm scratch arena (train_ops.ZeroArena +
gfr_bn_config), the job table of the one-launch operand packing (gfr_conv_tc_pack_job_fill) and the running-statistics
recorder used by the single discriminator pass (TRAIN:619 / 641).  No kernel is launched."""
import ctypes

import pytest
import torch

from geomconsistentfr_b200 import _lib, train_ops


def test_bn_config_is_a_process_wide_switch_that_returns_the_previous_setting():
    lib = _lib.load()
    assert lib.gfr_bn_config(0) in (0, 1)
    assert lib.gfr_bn_config(1) == 0
    assert lib.gfr_bn_config(7) == 1          # other values only query
    assert lib.gfr_bn_config(0) == 1
    assert lib.gfr_bn_config(-1) == 0


def test_zero_arena_slices_alignment_high_water_and_overflow():
    lib = _lib.load()
    lib.gfr_bn_config(0)
    a = train_ops.ZeroArena(n_doubles=64)
    assert a.take(5) is None                                  # before begin(): no buffer, the caller allocates
    with a:
        assert lib.gfr_bn_config(-1) == 1                     # the library skips its memsets while the arena is active
        a.begin(torch.device("cpu"))
        s1 = a.take(5)
        s2 = a.take(9)
        assert s1.numel() == 5 and s2.numel() == 9 and s1.dtype == torch.float64
        assert (s2.data_ptr() - s1.data_ptr()) == 6 * 8       # 16-byte aligned slices: 5 doubles take 6
        assert a.high == 16
        s1.fill_(3.0); s2.fill_(4.0)
        big = a.take(100)                                     # does not fit: a zeroed temporary of the right size, the arena untouched
        assert big.numel() == 100 and float(big.abs().sum()) == 0.0 and a.off == 16
        a.begin(torch.device("cpu"))                          # next step: the used part is cleared, slices restart at 0
        t1 = a.take(5)
        assert t1.data_ptr() == s1.data_ptr() and float(t1.abs().sum()) == 0.0 and float(s2.abs().sum()) == 0.0
        assert train_ops._bn_scratch(7, torch.device("cpu")).data_ptr() == s1.data_ptr() + 6 * 8
    assert lib.gfr_bn_config(-1) == 0 and train_ops.ZeroArena.active is None
    assert train_ops._bn_scratch(7, torch.device("cpu")).numel() == 7      # no arena: an ordinary temporary


@pytest.mark.parametrize("Cin,Cout,NT,taps,prec,dgrad", [(16, 16, 16, 9, 4, 0), (155, 64, 64, 9, 4, 1), (256, 128, 128, 4, 4, 0), (32, 32, 32, 9, 3, 0)])
def test_pack_job_table_matches_the_single_layer_pack_sizes(Cin, Cout, NT, taps, prec, dgrad):
    lib = _lib.load()
    rec = lib.gfr_conv_tc_pack_job_size()
    assert rec > 0 and rec % 8 == 0
    buf = ctypes.create_string_buffer(2 * rec)
    w = torch.zeros(Cout * Cin * taps)
    O, I = (Cin, Cout) if dgrad else (Cout, Cin)
    n_floats = lib.gfr_conv_tc_pack_size_ex(I, O, NT, taps, prec)
    packed = torch.zeros(n_floats)
    n_blocks = lib.gfr_conv_tc_pack_job_fill(ctypes.c_void_p(ctypes.addressof(buf)), ctypes.c_void_p(w.data_ptr()), 0, dgrad, Cin, Cout, NT,
                                             taps, prec, ctypes.c_void_p(packed.data_ptr()), 11)
    elems = n_floats * 2 if prec == 4 else n_floats                       # bf16: two elements per float of the buffer
    assert n_blocks == (elems + 255) // 256
    assert buf.raw[:rec] != b"\0" * rec and buf.raw[rec:] == b"\0" * rec   # exactly one record written
    # errors: bad tile width / taps, null pointers
    assert lib.gfr_conv_tc_pack_job_fill(ctypes.c_void_p(ctypes.addressof(buf)), ctypes.c_void_p(w.data_ptr()), 0, dgrad, Cin, Cout, 48, taps,
                                         prec, ctypes.c_void_p(packed.data_ptr()), 0) < 0
    assert lib.gfr_conv_tc_pack_job_fill(ctypes.c_void_p(ctypes.addressof(buf)), None, 0, dgrad, Cin, Cout, NT, taps, prec,
                                         ctypes.c_void_p(packed.data_ptr()), 0) < 0
    assert lib.gfr_conv_tc_pack_weights_batch(None, 1, 1, None) < 0
    assert lib.gfr_conv_tc_pack_weights_batch(ctypes.c_void_p(ctypes.addressof(buf)), 0, 1, None) < 0


def test_pack_plan_serves_parameters_only_and_tracks_freshness():
    p = torch.nn.Parameter(torch.zeros(16, 16, 3, 3))
    other = torch.zeros(16, 16, 3, 3)
    plan = train_ops.PackPlan([p])
    assert plan.get(other, False, False, 16, 16, 16, 9, 4) is None        # a temporary has no stable address
    buf, filled = plan.get(p, False, False, 16, 16, 16, 9, 4)
    assert not filled and buf.numel() == _lib.load().gfr_conv_tc_pack_size_ex(16, 16, 16, 9, 4)
    buf2, filled2 = plan.get(p, False, False, 16, 16, 16, 9, 4)
    assert buf2.data_ptr() == buf.data_ptr() and not filled2             # registered, but no run() yet: the caller packs it itself
    assert plan.get(p, False, True, 16, 16, 16, 9, 4)[0].data_ptr() != buf.data_ptr()      # the data-gradient operand is its own job
    assert len(plan.jobs) == 2
    plan.fresh = True                                                     # (what run() sets) — still not in the device table
    assert plan.get(p, False, False, 16, 16, 16, 9, 4)[1] is False
    for job in plan.jobs.values():
        job[9] = True
    assert plan.get(p, False, False, 16, 16, 16, 9, 4)[1] is True
    plan.invalidate()                                                     # the optimiser moved the parameters
    assert plan.get(p, False, False, 16, 16, 16, 9, 4)[1] is False
    with plan:
        assert train_ops.PackPlan.active is plan
    assert train_ops.PackPlan.active is None


def test_running_update_recorder_nests_and_restores():
    assert train_ops._BN.recording is None
    with train_ops.record_running_updates() as outer:
        assert train_ops._BN.recording is outer.records
        with train_ops.record_running_updates() as inner:
            assert train_ops._BN.recording is inner.records
        assert train_ops._BN.recording is outer.records
    assert train_ops._BN.recording is None
    train_ops._BN.replay_running_updates([])                              # nothing recorded: no launch, no generation bump
