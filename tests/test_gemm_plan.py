"""CPU: the GEMM work-unit plan (mico_b200/csrc/gemm.cu:GemmPlan) -- host logic only, no device.

Every unit must appear exactly once, a round's units stay in that round (concurrent CTAs share A panels through L2) and the
planned makespan must sit within one tile of the ideal for the tower's N = 1408 GEMMs, where the strided schedule lost 8 %."""
import ctypes as C

import numpy as np


def _plan(num_tiles, num_n, bn, n_last, num_kb, slots, allow_split):
    from mico_b200 import _lib
    L = _lib.lib
    L.mico_gemm_plan.restype = C.c_int
    ks, rounds, ms = C.c_int(0), C.c_int(0), C.c_double(0)
    args = [C.c_int(num_tiles), C.c_int(num_n), C.c_int(bn), C.c_int(n_last), C.c_int(num_kb), C.c_int(slots), C.c_int(allow_split)]
    assert L.mico_gemm_plan(*args, C.byref(ks), C.byref(rounds), C.byref(ms), None, 0) == 0
    tab = (C.c_int * (slots * rounds.value))()
    assert L.mico_gemm_plan(*args, C.byref(ks), C.byref(rounds), C.byref(ms), tab, len(tab)) == 0
    return ks.value, rounds.value, ms.value, np.array(tab, dtype=np.int64).reshape(slots, rounds.value)


def _costs(num_tiles, num_n, bn, n_last, num_kb, ks, drain=4):
    kb_per = -(-num_kb // ks)
    u = np.arange(num_tiles * ks)
    w = np.where((u % num_tiles) % num_n == num_n - 1, n_last, bn)
    kbc = np.clip(num_kb - (u // num_tiles) * kb_per, 0, kb_per)
    return w * (kbc + drain)


def _check(num_tiles, num_n, bn, n_last, num_kb, slots, allow_split):
    ks, rounds, ms, tab = _plan(num_tiles, num_n, bn, n_last, num_kb, slots, allow_split)
    cost = _costs(num_tiles, num_n, bn, n_last, num_kb, ks)
    n_units = num_tiles * ks
    used = tab[tab >= 0]
    assert sorted(used.tolist()) == list(range(n_units))              # every unit exactly once
    for s in range(slots):                                            # lists are compact, in round order
        row = tab[s]
        k = int((row >= 0).sum())
        assert (row[:k] >= 0).all() and (row[k:] == -1).all()
        assert (np.diff(row[:k] // slots) > 0).all()                  # one unit per round, rounds ascending
    load = np.array([cost[tab[s][tab[s] >= 0]].sum() for s in range(slots)])
    assert abs(load.max() - ms) < 1e-6 * ms
    strided = np.array([cost[s::slots].sum() for s in range(slots)]).max()
    return ks, load, strided, cost


def test_fc2_forward_at_omni_rows_is_balanced():
    # M = 197 376 rows as 771 pair groups, N = 1408 as five 256-wide tiles + one 128-wide, K = 6144 (96 blocks), 74 SM pairs
    ks, load, strided, cost = _check(771 * 6, 6, 256, 128, 96, 74, 0)
    assert ks == 1
    ideal = cost.sum() / 74
    assert load.max() <= ideal + cost.max() + 1e-9
    assert strided >= 1.07 * load.max()                               # what the strided schedule left on the table
    assert load.min() >= load.max() - cost.max()


def test_uniform_shapes_keep_the_strided_schedule():
    ks, rounds, ms, tab = _plan(65 * 24, 24, 256, 256, 22, 74, 0)     # N = 6144: every tile is full width
    exp = np.full((74, rounds), -1)
    for u in range(65 * 24):
        exp[u % 74, u // 74] = u
    assert (tab == exp).all()


def test_weight_gradient_split_is_chosen_by_makespan():
    # fc1 weight gradient of the omni tower pass: M = 6144 (24 pair groups), N = 1408, K = 197 376 tokens (3084 blocks)
    ks, load, strided, cost = _check(24 * 6, 6, 256, 128, 3084, 74, 1)
    assert ks in (2, 4)
    one = _costs(24 * 6, 6, 256, 128, 3084, 1)
    two_rounds = 2 * one.max()                                        # the unsplit launch: two full-K rounds
    assert load.max() <= 0.95 * two_rounds
    # qkv weight gradient: M = 4224 (17 pair groups, the last half empty), N = 1408
    ks, load, strided, cost = _check(17 * 6, 6, 256, 128, 3084, 74, 1)
    assert ks >= 2 and load.max() <= strided + 1e-9


def test_ragged_k_parts_and_small_launches():
    for args in [(3 * 2, 2, 128, 32, 70, 148, 1), (5, 1, 256, 256, 33, 74, 1), (7 * 3, 3, 176, 176, 12, 148, 0), (1, 1, 64, 16, 1, 148, 0)]:
        _check(*args)


def test_plan_properties_over_random_launch_shapes():
    """Property test (hypothesis): for any launch geometry the plan is a partition of the units into compact per-slot lists with
    one unit per round, its makespan is what the table says, never worse than the strided schedule's, and within one unit of the
    mean load when a round-robin deal could be balanced at all."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True, database=None)
    @given(st.integers(1, 40), st.integers(1, 7), st.sampled_from([64, 128, 176, 256]), st.integers(1, 16),
           st.integers(1, 300), st.sampled_from([74, 148, 70, 3]), st.booleans())
    def prop(mg, nn, bn, last16, num_kb, slots, split):
        n_last = min(bn, 16 * last16)
        ks, load, strided, cost = _check(mg * nn, nn, bn, n_last, num_kb, slots, int(split))
        assert ks in (1, 2, 4) and (split or ks == 1)
        assert load.max() <= strided + 1e-9
        assert load.max() <= cost.sum() / slots + 2 * cost.max() + 1e-9 or len(cost) < slots

    prop()
