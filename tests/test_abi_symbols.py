"""The C-ABI library loads and exports every symbol include/pdeb200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

from pde_controller_b200 import abi, build, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    return ctypes.CDLL(build.build())


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pdeb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdeb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_python_mirror_agree():
    assert declared_symbols() == sorted(abi.EXPORTS)


def test_every_declared_symbol_is_exported(so):
    for name in declared_symbols():
        assert hasattr(so, name), name


def test_every_pointer_taking_export_has_argtypes():
    """ctypes passes an undeclared argument as a C int: a 64-bit pointer would be truncated (a crash on the GPU box)."""
    L = native.lib()
    no_args = {"pdeb_version", "pdeb_device_count", "pdeb_last_error", "pdeb_cache_clear", "pdeb_controls_pool_clear"}
    for name in abi.EXPORTS:
        if name in no_args:
            continue
        assert getattr(L, name).argtypes, name


def test_version_and_defaults(so):
    assert so.pdeb_version() == abi.ABI_VERSION
    c = abi.PdebConventions()
    so.pdeb_default_conventions(ctypes.byref(c))
    d = abi.default_conventions()
    for f, _ in abi.PdebConventions._fields_:
        assert getattr(c, f) == getattr(d, f), f


def test_struct_sizes_match_header():
    # 8 int32 + 4 double; 2 int64 + 4 int32 + ptr + int64 + 28 ptr
    assert ctypes.sizeof(abi.PdebConventions) == 8 * 4 + 4 * 8
    assert ctypes.sizeof(abi.PdebBatch) == 2 * 8 + 4 * 4 + 8 + 8 + 28 * 8
    assert ctypes.sizeof(abi.PdebTiming) == 4 * 4 + 2 * 4
    # int64 + 4 int32 + 3 int64 + ptr + int64 + 28 ptr
    assert ctypes.sizeof(abi.PdebPackedBatch) == 8 + 4 * 4 + 3 * 8 + 8 + 8 + 28 * 8 + 24


def test_packed_transfer_form_round_trip():
    """PackedBatch (ragged lists, narrow counts) carries exactly the entries a batch uses: pack -> numpy unpack."""
    import numpy as np
    from pde_controller_b200 import PackedBatch, synth
    for wl, uk, n in (("heat", True, 1000), ("wave", False, 257), ("mixed", False, 2049), ("heat", False, 1)):
        b = synth.make_batch(n, wl, seed=3, uniform_knots=uk)
        b.n_knots[: min(5, n)] = 0                       # empty lists
        pk = PackedBatch(b.n, b.kmax, b.pmax, b.tmax, uniform_knots=uk).pack(b)
        assert pk.nbytes < b.slab_bytes and pk.nbytes <= PackedBatch.capacity_for(b.n, b.kmax, b.pmax, b.tmax, uk)
        u = pk.unpack()
        cnts = {"knot": b.n_knots, "pred": b.n_pred, "tree": b.n_tree}
        for name in b._names:
            a, c = getattr(b, name), getattr(u, name)
            if a.ndim == 1:
                assert np.array_equal(a, c), name
            else:
                live = np.arange(a.shape[0])[:, None] < cnts[name[:4]][None, :]
                assert np.array_equal(a[:, : b.n][live], c[:, : b.n][live]), name
        c = pk.as_c()
        assert (c.total_knots, c.total_pred, c.total_tree) == (int(b.n_knots.sum()), int(b.n_pred.sum()), int(b.n_tree.sum()))
        if n > 1:
            # the synthetic sets have integer lengths / interval ends (stored as float) and, per kind, one time step
            assert pk.enc["dt"] == (abi.ENC_F64 if wl == "mixed" else abi.ENC_ONE)
            assert pk.enc["L"] == abi.ENC_F32 and pk.enc["pred_xlo"] == abi.ENC_F32
            assert pk.enc["pred_m"] == abi.ENC_F64 and pk.enc["knot_y"] == abi.ENC_F64


def test_packed_encodings_are_lossless_bit_for_bit():
    """-0.0, infinities, NaNs and values one ulp away from a float must come back with the same BITS."""
    import numpy as np
    from pde_controller_b200 import PackedBatch, synth
    b = synth.make_batch(64, "heat", seed=4)
    b.g0[:] = 300.0
    b.g0[3] = -0.0                                       # not equal in bits to 0.0: no "one value" / stays exact as float
    b.L[5] = np.nextafter(100.0, 200.0)                  # not a float: the whole array stays f64
    b.x_a[:] = 0.0
    b.x_a[7] = -0.0                                      # all "equal" numerically, different bits
    b.pred_b[0, 9] = np.inf
    b.knot_y[0, 11] = np.nan
    pk = PackedBatch(b.n, b.kmax, b.pmax, b.tmax).pack(b)
    assert pk.enc["L"] == abi.ENC_F64 and pk.enc["g0"] == abi.ENC_F32 and pk.enc["x_a"] == abi.ENC_F32
    u = pk.unpack()
    for name in ("g0", "L", "x_a"):
        assert np.array_equal(getattr(u, name).view(np.uint64), getattr(b, name).view(np.uint64)), name
    live = np.arange(b.pmax)[:, None] < b.n_pred[None, :]
    assert np.array_equal(u.pred_b[live].view(np.uint64), b.pred_b[live].view(np.uint64))
    livek = np.arange(b.kmax)[:, None] < b.n_knots[None, :]
    assert np.array_equal(u.knot_y[livek].view(np.uint64), b.knot_y[livek].view(np.uint64))


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    if native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from pde_controller_b200 import synth
    b = synth.make_batch(4, "heat")
    with pytest.raises(native.NativeError):
        native.evaluate_batch(b)
    with pytest.raises(native.NativeError):
        native.ControlsEvaluator(b, [0.0, 1.0], capacity=8)



def test_native_packer_equals_numpy_packer():
    """pdeb_pack_batch (host-side C, no device involved) and the numpy packer produce the same block: sizes, storage
    tags, totals and decoded content, including the bit-level corner cases."""
    import numpy as np
    from pde_controller_b200 import PackedBatch, synth
    L = native.lib()
    cases = [synth.make_batch(n, wl, seed=3, uniform_knots=uk)
             for wl, uk, n in (("heat", True, 4096), ("wave", False, 3001), ("mixed", False, 2049), ("heat", False, 1))]
    odd = synth.make_batch(64, "heat", seed=4)
    odd.g0[3] = -0.0
    odd.L[5] = np.nextafter(100.0, 200.0)
    odd.x_a[:] = 0.0
    odd.x_a[7] = -0.0
    odd.pred_b[0, 9] = np.inf
    odd.knot_y[0, 11] = np.nan
    odd.E_in[:] = 1.0e6
    cases.append(odd)
    for b in cases:
        b.n_knots[: min(5, b.n)] = 0
        a = PackedBatch(b.n, b.kmax, b.pmax, b.tmax, uniform_knots=b.uniform_knots).pack(b, use_native=False)
        c = PackedBatch(b.n, b.kmax, b.pmax, b.tmax, uniform_knots=b.uniform_knots).pack(b, use_native=True)
        assert (a.nbytes, a.totals, a.enc) == (c.nbytes, c.totals, c.enc)
        ua, uc = a.unpack(), c.unpack()
        for name in b._names:
            if name == "knot_t" and b.uniform_knots:
                continue
            assert np.array_equal(getattr(ua, name).view(np.uint8), getattr(uc, name).view(np.uint8)), name
        assert L.pdeb_packed_capacity(b.n, b.kmax, b.pmax, b.tmax, 0 if b.uniform_knots else 1) <= \
            PackedBatch.capacity_for(b.n, b.kmax, b.pmax, b.tmax, b.uniform_knots)
    big = synth.make_batch(8, "heat", seed=1)
    big.n_elem[2] = 70000                                  # does not fit u16
    with pytest.raises(ValueError):
        PackedBatch(big.n, big.kmax, big.pmax, big.tmax).pack(big, use_native=True)


def test_cost_balanced_shard_bounds_match_the_python_rule():
    """pdeb_shard_bounds (host-only): contiguous shards of equal cost (steps x rows, x 1.5 for waves: SURVEY.md 8e),
    the same cut points as sharding.cost_balanced_bounds."""
    import numpy as np
    from pde_controller_b200 import sharding, synth
    b = synth.make_batch(6000, "mixed", seed=2)            # first half heat, second half waves (3x the cost each)
    for ndev in (1, 2, 3, 8):
        got = native.shard_bounds(b, ndev)
        want = sharding.cost_balanced_bounds(sharding.batch_cost(b), ndev)
        assert got[0] == 0 and got[-1] == b.n and (np.diff(got) >= 0).all()
        assert [(int(got[r]), int(got[r + 1])) for r in range(ndev)] == want
        cost = sharding.batch_cost(b)
        per = np.array([cost[got[r]:got[r + 1]].sum() for r in range(ndev)])
        assert per.max() <= per.mean() * 1.02 + cost.max()
    assert list(native.shard_bounds(b, 2)) != [0, b.n // 2, b.n]      # by cost, not by count
