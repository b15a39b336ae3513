"""ctypes mirror of include/pdeb200.h (data layout only, no behaviour)."""
import ctypes as C

ABI_VERSION = 1

KIND_HEAT, KIND_WAVE = 0, 1
REL_GT, REL_LT, REL_EQ = 0, 1, 2
TOP_G, TOP_F = 0, 1
TREE_LEAF, TREE_AND, TREE_OR = 0, 1, 2
OK, ST_PARSE, ST_NONFINITE, ST_BADPARAM = 0, 1, 2, 3
FAILED_ROBUSTNESS = -100.0
MAX_PRED = 8
MAX_TREE = 2 * MAX_PRED - 1
MAX_ELEM = 1054
MAX_STEPS = 1 << 22

REL_CODE = {">": REL_GT, "<": REL_LT, "=": REL_EQ}
REL_TEXT = {v: k for k, v in REL_CODE.items()}
TOP_CODE = {"G": TOP_G, "F": TOP_F}
TOP_TEXT = {v: k for k, v in TOP_CODE.items()}


class PdebConventions(C.Structure):
    _fields_ = [
        ("material_sampling", C.c_int32),
        ("mass_heat", C.c_int32),
        ("mass_wave", C.c_int32),
        ("steps_rule", C.c_int32),
        ("window_rule", C.c_int32),
        ("region_rule", C.c_int32),
        ("eq_rule", C.c_int32),
        ("reserved0", C.c_int32),
        ("eq_eps", C.c_double),
        ("theta", C.c_double),
        ("newmark_beta", C.c_double),
        ("newmark_gamma", C.c_double),
    ]


def default_conventions(**overrides):
    """The "v0" convention set of SURVEY.md §8c."""
    c = PdebConventions()
    c.material_sampling = 0
    c.mass_heat = 0
    c.mass_wave = 1
    c.steps_rule = 0
    c.window_rule = 0
    c.region_rule = 0
    c.eq_rule = 0
    c.eq_eps = 0.0
    c.theta = 0.5
    c.newmark_beta = 0.25
    c.newmark_gamma = 0.5
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError(f"unknown convention flag {k!r}")
        setattr(c, k, v)
    return c


_P = C.c_void_p


class PdebBatch(C.Structure):
    _fields_ = [
        ("n", C.c_int64),
        ("stride", C.c_int64),
        ("kmax", C.c_int32),
        ("pmax", C.c_int32),
        ("tmax", C.c_int32),
        ("reserved0", C.c_int32),
        ("slab_base", _P),
        ("slab_bytes", C.c_int64),
        ("kind", _P),
        ("n_elem", _P),
        ("L", _P),
        ("dt", _P),
        ("T", _P),
        ("g0", _P),
        ("x_a", _P),
        ("x_b", _P),
        ("rho_in", _P),
        ("rho_out", _P),
        ("E_in", _P),
        ("E_out", _P),
        ("n_knots", _P),
        ("knot_t", _P),
        ("knot_y", _P),
        ("n_pred", _P),
        ("pred_xlo", _P),
        ("pred_xhi", _P),
        ("pred_m", _P),
        ("pred_b", _P),
        ("pred_tlo", _P),
        ("pred_thi", _P),
        ("pred_rel", _P),
        ("pred_top", _P),
        ("n_tree", _P),
        ("tree_op", _P),
        ("tree_lhs", _P),
        ("tree_rhs", _P),
    ]


class PdebPackedBatch(C.Structure):
    """Packed transfer form (include/pdeb200.h): ragged lists, narrow counts, one contiguous block."""
    _fields_ = [
        ("n", C.c_int64),
        ("kmax", C.c_int32),
        ("pmax", C.c_int32),
        ("tmax", C.c_int32),
        ("reserved0", C.c_int32),
        ("total_knots", C.c_int64),
        ("total_pred", C.c_int64),
        ("total_tree", C.c_int64),
        ("slab_base", _P),
        ("slab_bytes", C.c_int64),
        ("kind", _P), ("n_elem", _P),
        ("L", _P), ("dt", _P), ("T", _P), ("g0", _P), ("x_a", _P), ("x_b", _P),
        ("rho_in", _P), ("rho_out", _P), ("E_in", _P), ("E_out", _P),
        ("n_knots", _P), ("n_pred", _P), ("n_tree", _P),
        ("knot_t", _P), ("knot_y", _P),
        ("pred_xlo", _P), ("pred_xhi", _P), ("pred_m", _P), ("pred_b", _P), ("pred_tlo", _P), ("pred_thi", _P),
        ("pred_rel", _P), ("pred_top", _P),
        ("tree_op", _P), ("tree_lhs", _P), ("tree_rhs", _P),
        ("enc", C.c_uint8 * 24),
    ]


# lossless encodings of the real-valued arrays of the packed form, and their slots in PdebPackedBatch.enc
ENC_F64, ENC_F32, ENC_ONE = 0, 1, 2
PK_INDEX = {name: i for i, name in enumerate(
    ["L", "dt", "T", "g0", "x_a", "x_b", "rho_in", "rho_out", "E_in", "E_out", "knot_t", "knot_y",
     "pred_xlo", "pred_xhi", "pred_m", "pred_b", "pred_tlo", "pred_thi"])}


class PdebTiming(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float),
        ("kernel_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("total_ms", C.c_float),
        ("launches", C.c_int32),
        ("reserved0", C.c_int32),
    ]


# every symbol include/pdeb200.h declares (checked by tests/test_abi_symbols.py)
class PdebIouClause(C.Structure):
    """One clause of an IoU comparison (include/pdeb200.h)."""
    _fields_ = [("t0", C.c_double), ("t1", C.c_double), ("p0", C.c_double), ("p1", C.c_double), ("coef", C.c_double),
                ("c", C.c_double), ("thr", C.c_double), ("item", C.c_int32), ("kind", C.c_uint8), ("neg", C.c_uint8),
                ("op", C.c_uint8), ("backwards", C.c_uint8)]


IOU_MUL, IOU_DIV = 0, 1
IOU_OP = {"<=": 0, "<": 1, ">=": 2, ">": 3, "=": 4, "!=": 5}

EXPORTS = [
    "pdeb_version",
    "pdeb_device_count",
    "pdeb_device_pci_bus_id",
    "pdeb_last_error",
    "pdeb_default_conventions",
    "pdeb_eval_batch",
    "pdeb_eval_batch_multi",
    "pdeb_shard_bounds",
    "pdeb_parse_programs",
    "pdeb_iou_grids",
    "pdeb_cache_clear",
    "pdeb_resident_create",
    "pdeb_resident_upload",
    "pdeb_resident_upload_packed",
    "pdeb_packed_capacity",
    "pdeb_pack_batch",
    "pdeb_resident_run",
    "pdeb_resident_eval",
    "pdeb_resident_set_knots",
    "pdeb_resident_plan_order",
    "pdeb_controls_create",
    "pdeb_controls_eval",
    "pdeb_controls_destroy",
    "pdeb_controls_respec",
    "pdeb_controls_search",
    "pdeb_controls_search_batch",
    "pdeb_controls_pool_clear",
    "pdeb_controls_fetch_knots",
    "pdeb_resident_fetch",
    "pdeb_resident_fetch_async",
    "pdeb_stream_create",
    "pdeb_stream_destroy",
    "pdeb_stream_sync",
    "pdeb_resident_device_ptrs",
    "pdeb_resident_destroy",
    "pdeb_host_alloc",
    "pdeb_host_free",
]
