"""ctypes binding of libpdeb200.so (the C ABI of include/pdeb200.h).

There is no CPU fallback: if the library is missing or no CUDA device is visible, the
evaluation entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import abi
from .problem import ProblemBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PDEB_LIB lets the kernel-variant experiments in tools/ load an alternative build of the same ABI
SO_PATH = os.environ.get("PDEB_LIB") or os.path.join(_HERE, "libpdeb200.so")
_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load libpdeb200.so (built in-tree by pde_controller_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise NativeError(f"{SO_PATH} is missing: run `python -m pde_controller_b200.build` "
                              "(there is no CPU fallback for this path)")
        L = C.CDLL(SO_PATH)
        L.pdeb_version.restype = C.c_int
        L.pdeb_device_count.restype = C.c_int
        L.pdeb_last_error.restype = C.c_char_p
        L.pdeb_device_pci_bus_id.restype = C.c_int
        L.pdeb_device_pci_bus_id.argtypes = [C.c_int, C.c_char_p, C.c_int]
        L.pdeb_default_conventions.argtypes = [C.c_void_p]
        L.pdeb_host_alloc.restype = C.c_void_p
        L.pdeb_host_alloc.argtypes = [C.c_int64]
        L.pdeb_host_free.argtypes = [C.c_void_p]
        L.pdeb_eval_batch.restype = C.c_int
        L.pdeb_eval_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.pdeb_eval_batch_multi.restype = C.c_int
        L.pdeb_eval_batch_multi.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_shard_bounds.restype = C.c_int
        L.pdeb_shard_bounds.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.pdeb_parse_programs.restype = C.c_int
        L.pdeb_parse_programs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.pdeb_iou_grids.restype = C.c_int
        L.pdeb_iou_grids.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.pdeb_cache_clear.restype = None
        L.pdeb_cache_clear.argtypes = []
        L.pdeb_resident_create.restype = C.c_int
        L.pdeb_resident_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.pdeb_resident_upload.restype = C.c_int
        L.pdeb_resident_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_packed_capacity.restype = C.c_int64
        L.pdeb_packed_capacity.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.pdeb_pack_batch.restype = C.c_int
        L.pdeb_pack_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.pdeb_resident_upload_packed.restype = C.c_int
        L.pdeb_resident_upload_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_resident_run.restype = C.c_int
        L.pdeb_resident_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.pdeb_resident_eval.restype = C.c_int
        L.pdeb_resident_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_resident_set_knots.restype = C.c_int
        L.pdeb_resident_set_knots.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_resident_fetch.restype = C.c_int
        L.pdeb_resident_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_resident_fetch_async.restype = C.c_int
        L.pdeb_resident_fetch_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_stream_create.restype = C.c_void_p
        L.pdeb_stream_create.argtypes = [C.c_int]
        L.pdeb_stream_destroy.argtypes = [C.c_void_p]
        L.pdeb_stream_sync.restype = C.c_int
        L.pdeb_stream_sync.argtypes = [C.c_void_p]
        L.pdeb_resident_plan_order.restype = C.c_int
        L.pdeb_resident_plan_order.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_resident_device_ptrs.restype = C.c_int
        L.pdeb_resident_device_ptrs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_resident_destroy.argtypes = [C.c_void_p]
        L.pdeb_controls_create.restype = C.c_int
        L.pdeb_controls_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.pdeb_controls_eval.restype = C.c_int
        L.pdeb_controls_eval.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.pdeb_controls_destroy.argtypes = [C.c_void_p]
        L.pdeb_controls_respec.restype = C.c_int
        L.pdeb_controls_respec.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.pdeb_controls_search.restype = C.c_int
        L.pdeb_controls_search.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        L.pdeb_controls_search_batch.restype = C.c_int
        L.pdeb_controls_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_void_p,
                                                 C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p]
        L.pdeb_controls_pool_clear.restype = None
        L.pdeb_controls_pool_clear.argtypes = []
        L.pdeb_controls_fetch_knots.restype = C.c_int
        L.pdeb_controls_fetch_knots.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        if L.pdeb_version() != abi.ABI_VERSION:
            raise NativeError("libpdeb200.so ABI version mismatch: rebuild it")
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {lib().pdeb_last_error().decode()}")


def device_count() -> int:
    return int(lib().pdeb_device_count())


def device_pci_bus_id(device: int) -> str:
    buf = C.create_string_buffer(32)
    _check(lib().pdeb_device_pci_bus_id(int(device), buf, 32), "pdeb_device_pci_bus_id")
    return buf.value.decode().lower()


def _cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _nvml_cpu_affinity(bus_id: str) -> set:
    """The driver's ideal CPU set of the device at `bus_id` (what `nvidia-smi topo -m` prints), empty if NVML is not there."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(os.cpu_count() or 1, 1) + 63) // 64)
            return {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return set()


def bind_host_near(device: int, sysfs: str = "/sys") -> dict:
    """Run the calling thread (and the threads it starts later) on the CPUs of the NUMA node `device` hangs off.

    Page-locked buffers allocated afterwards (`pdeb_host_alloc`, the staging mirrors of the one-shot calls) are placed on
    that node by the kernel's local allocation policy, so host -> device copies do not cross the socket interconnect: on a box with
    one process per GPU that is what keeps eight concurrent uploads from sharing one socket's links.  Returns what was
    done: {"bus_id", "node", "cpus", "bound"}; on a box without NUMA information (node -1, containers that hide sysfs)
    nothing changes and "bound" is False.  Undo with os.sched_setaffinity(0, previous).
    """
    info = {"bus_id": None, "node": None, "cpus": 0, "bound": False}
    try:
        bus = device_pci_bus_id(device)
        info["bus_id"] = bus
        with open(f"{sysfs}/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        info["node"] = node
        if node < 0:
            cpus = _nvml_cpu_affinity(bus) & os.sched_getaffinity(0)     # virtual machines: sysfs says -1, the driver may still know
            if not cpus or cpus == os.sched_getaffinity(0):
                return info
        else:
            with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as fh:
                cpus = _cpulist(fh.read()) & os.sched_getaffinity(0)
        info["cpus"] = len(cpus)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except (OSError, ValueError, NativeError):
        pass
    return info


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data if a is not None else None


def _conv_ptr(conv):
    return C.addressof(conv) if conv is not None else None


class PinnedBuffer:
    """Page-locked host memory exposing the buffer protocol through numpy."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().pdeb_host_alloc(self.nbytes)
        if not self.ptr:
            raise NativeError("pinned allocation failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self.array = None
            lib().pdeb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_batch(n, kmax, pmax, tmax=None, uniform_knots=False) -> ProblemBatch:
    """An empty ProblemBatch whose slab is page-locked (fill it, then evaluate it every step)."""
    buf = PinnedBuffer(ProblemBatch.nbytes_for(n, kmax, pmax, tmax, uniform_knots))
    b = ProblemBatch(n, kmax, pmax, tmax, buffer=buf.array, uniform_knots=uniform_knots)
    b._pinned = buf
    return b


def pinned_packed(n, kmax, pmax, tmax=None, uniform_knots=False):
    """An empty PackedBatch in page-locked memory (pack() into it, then upload_packed it)."""
    from .problem import PackedBatch
    buf = PinnedBuffer(PackedBatch.capacity_for(n, kmax, pmax, tmax, uniform_knots) + 64)
    p = PackedBatch(n, kmax, pmax, tmax, buffer=buf.array, uniform_knots=uniform_knots)
    p._pinned = buf
    return p


def evaluate_batch(batch: ProblemBatch, device: int = 0, precision: int = 64, conv=None, want_clauses: bool = False,
                   stream: int = 0, timing: bool = False):
    """One-shot: H2D, kernels, D2H.  -> dict(robustness, status[, clause][, timing])."""
    L = lib()
    cb = batch.as_c()
    rob = np.empty(batch.n, dtype=np.float64)
    st = np.empty(batch.n, dtype=np.int32)
    cl = np.empty((batch.pmax, batch.stride), dtype=np.float64) if want_clauses else None
    tm = abi.PdebTiming() if timing else None
    rc = L.pdeb_eval_batch(C.addressof(cb), device, precision, _conv_ptr(conv), _ptr(rob), _ptr(cl), _ptr(st),
                           stream or None, C.addressof(tm) if tm is not None else None)
    _check(rc, "pdeb_eval_batch")
    out = {"robustness": rob, "status": st}
    if want_clauses:
        out["clause"] = cl[:, : batch.n]
    if timing:
        out["timing"] = tm
    return out


def evaluate_batch_multi(batch: ProblemBatch, ndev: int, devices=None, precision: int = 64, conv=None,
                         want_clauses: bool = False):
    """Contiguous shards over ndev devices of this box (one host thread + stream per device)."""
    L = lib()
    cb = batch.as_c()
    rob = np.empty(batch.n, dtype=np.float64)
    st = np.empty(batch.n, dtype=np.int32)
    cl = np.empty((batch.pmax, batch.stride), dtype=np.float64) if want_clauses else None
    tms = (abi.PdebTiming * ndev)()
    devs = (C.c_int * ndev)(*devices) if devices is not None else None
    rc = L.pdeb_eval_batch_multi(C.addressof(cb), ndev, devs, precision, _conv_ptr(conv), _ptr(rob), _ptr(cl), _ptr(st),
                                 C.addressof(tms))
    _check(rc, "pdeb_eval_batch_multi")
    out = {"robustness": rob, "status": st, "timing": list(tms)}
    if want_clauses:
        out["clause"] = cl[:, : batch.n]
    return out


def shard_bounds(batch: ProblemBatch, ndev: int, conv=None) -> np.ndarray:
    """Cut points of the cost-balanced contiguous shards pdeb_eval_batch_multi uses (host-side rule, no device)."""
    cb = batch.as_c()
    out = np.zeros(ndev + 1, dtype=np.int64)
    _check(lib().pdeb_shard_bounds(C.addressof(cb), ndev, _conv_ptr(conv), _ptr(out)), "pdeb_shard_bounds")
    return out


def controls_pool_clear():
    """Free the controls handles finished search sweeps have left for the next one."""
    lib().pdeb_controls_pool_clear()


def cache_clear():
    """Free the resident images one-shot calls keep per device and shape."""
    lib().pdeb_cache_clear()


class ResidentBatch:
    """A batch kept in HBM: upload once (or refill), evaluate many times."""

    def __init__(self, batch: ProblemBatch, device: int = 0, stream: int = 0):
        self._L = lib()
        self.batch = batch
        self.device = device
        self._h = C.c_void_p()
        cb = batch.as_c()
        _check(self._L.pdeb_resident_create(C.addressof(cb), device, stream or None, C.addressof(self._h)),
               "pdeb_resident_create")

    def upload(self, batch: Optional[ProblemBatch] = None, stream: int = 0):
        cb = (batch or self.batch).as_c()
        _check(self._L.pdeb_resident_upload(self._h, C.addressof(cb), stream or None), "pdeb_resident_upload")

    def upload_packed(self, packed, stream: int = 0):
        """Refill the image from the packed transfer form (problem.PackedBatch): one smaller copy, unpacked on the device."""
        cp = packed.as_c()
        _check(self._L.pdeb_resident_upload_packed(self._h, C.addressof(cp), stream or None), "pdeb_resident_upload_packed")

    def run(self, batch: Optional[ProblemBatch] = None, precision: int = 64, conv=None, want_clauses: bool = False,
            stream: int = 0, out=None, timing: bool = False):
        """Upload + evaluate + fetch in one pipelined pass (the end-to-end path for a fresh batch of this shape)."""
        b = batch or self.batch
        cb = b.as_c()
        rob = out["robustness"] if out else np.empty(b.n, dtype=np.float64)
        st = out["status"] if out else np.empty(b.n, dtype=np.int32)
        cl = np.empty((b.pmax, b.stride), dtype=np.float64) if want_clauses else None
        tm = abi.PdebTiming() if timing else None
        _check(self._L.pdeb_resident_run(self._h, C.addressof(cb), precision, _conv_ptr(conv), _ptr(rob), _ptr(cl), _ptr(st),
                                         stream or None, C.addressof(tm) if tm is not None else None), "pdeb_resident_run")
        res = {"robustness": rob, "status": st}
        if want_clauses:
            res["clause"] = cl[:, : b.n]
        if timing:
            res["timing"] = tm
        return res

    def eval(self, precision: int = 64, conv=None, stream: int = 0, timing: bool = False):
        tm = abi.PdebTiming() if timing else None
        _check(self._L.pdeb_resident_eval(self._h, precision, _conv_ptr(conv), stream or None,
                                          C.addressof(tm) if tm is not None else None), "pdeb_resident_eval")
        return tm

    def set_knots(self, knot_t: np.ndarray, knot_y: np.ndarray, n_knots: np.ndarray, stream: int = 0):
        _check(self._L.pdeb_resident_set_knots(self._h, _ptr(knot_t), _ptr(knot_y), _ptr(n_knots), stream or None),
               "pdeb_resident_set_knots")

    def fetch(self, want_clauses: bool = False, stream: int = 0, out=None):
        b = self.batch
        rob = out["robustness"] if out else np.empty(b.n, dtype=np.float64)
        st = out["status"] if out else np.empty(b.n, dtype=np.int32)
        cl = np.empty((b.pmax, b.stride), dtype=np.float64) if want_clauses else None
        _check(self._L.pdeb_resident_fetch(self._h, _ptr(rob), _ptr(cl), _ptr(st), stream or None), "pdeb_resident_fetch")
        res = {"robustness": rob, "status": st}
        if want_clauses:
            res["clause"] = cl[:, : b.n]
        return res

    def fetch_async(self, out, stream: int = 0):
        """Enqueue the D2H copies of robustness + status into `out` (page-locked arrays); no synchronisation."""
        _check(self._L.pdeb_resident_fetch_async(self._h, _ptr(out["robustness"]), None, _ptr(out["status"]), stream or None),
               "pdeb_resident_fetch_async")

    def plan_order(self, stream: int = 0) -> np.ndarray:
        """Sample ids in the order the kernels walk them (as formed under the conventions of the last eval)."""
        order = np.empty(self.batch.n, dtype=np.int32)
        _check(self._L.pdeb_resident_plan_order(self._h, _ptr(order), stream or None), "pdeb_resident_plan_order")
        return order

    def device_ptrs(self):
        r, c, s = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self._L.pdeb_resident_device_ptrs(self._h, C.addressof(r), C.addressof(c), C.addressof(s)),
               "pdeb_resident_device_ptrs")
        return r.value, c.value, s.value

    def close(self):
        if self._h:
            self._L.pdeb_resident_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BlockPipeline:
    """Streams batches through the GPU as fixed-size blocks, multi-buffered on `max(2, depth)` streams.

    Every block has its own page-locked host slab (one H2D copy), its own resident image and its own slice of
    page-locked result arrays; the upload of one block overlaps the kernels of the previous one and the download
    of the one before.  `depth` generations of buffers let consecutive batches overlap as well: submit(g) enqueues
    generation g and returns at once, wait(g) blocks until its results are in host memory.  This is the end-to-end
    route for the 2M-sample sweep: fill(batch, g), submit(g), ..., wait(g), read robustness[g] / status[g].
    """

    def __init__(self, n: int, kmax: int, pmax: int, tmax=None, block: int = 65536, device: int = 0, depth: int = 2,
                 uniform_knots: bool = False, packed: bool = False):
        self._L = lib()
        self.n, self.device, self.depth = int(n), device, int(depth)
        self.bounds = [(lo, min(lo + block, n)) for lo in range(0, n, block)]
        # one stream per generation (at least two: the blocks of one generation alternate between streams), so that
        # waiting for generation g never waits for work submitted after it
        self.streams = [self._L.pdeb_stream_create(device) for _ in range(max(2, self.depth))]
        if not all(self.streams):
            raise NativeError("stream creation failed")
        self.packed = bool(packed)       # blocks cross PCIe in the packed transfer form (ragged lists, narrow counts)
        self.packs = []
        self.blocks, self.residents, self.robustness, self.status, self._outs, self._pins = [], [], [], [], [], []
        slot = 0
        for g in range(self.depth):
            rob, st = PinnedBuffer(8 * max(n, 1)), PinnedBuffer(4 * max(n, 1))
            self._pins += [rob, st]
            self.robustness.append(rob.array.view(np.float64)[:n])
            self.status.append(st.array.view(np.int32)[:n])
            blks = [pinned_batch(hi - lo, kmax, pmax, tmax, uniform_knots) for lo, hi in self.bounds]
            self.blocks.append(blks)
            self.packs.append([pinned_packed(hi - lo, kmax, pmax, tmax, uniform_knots) for lo, hi in self.bounds]
                              if self.packed else [None] * len(self.bounds))
            self._outs.append([{"robustness": self.robustness[g][lo:hi], "status": self.status[g][lo:hi]}
                               for lo, hi in self.bounds])
            res = []
            for b in blks:
                si = slot % len(self.streams)
                res.append((ResidentBatch(b, device=device, stream=self.streams[si]), si))
                slot += 1
            self.residents.append(res)
        self._sync_all()

    def fill(self, batch: ProblemBatch, gen: int = 0):
        """Copy a whole batch into the blocks of one generation (host side)."""
        for (lo, hi), blk, pk in zip(self.bounds, self.blocks[gen], self.packs[gen]):
            for name in blk._names:
                getattr(blk, name)[...] = getattr(batch, name)[..., lo:hi]
            if pk is not None:
                pk.pack(blk)

    def submit(self, gen: int = 0, precision: int = 64, conv=None, pack: bool = False):
        """Enqueue upload + plan + kernels + download of every block of generation `gen`; does not wait.
        pack (packed pipelines only): run the host-side producer of the packed form (PackedBatch.pack ->
        pdeb_pack_batch) on each block's plane-major batch first, i.e. inside whatever region the caller is timing."""
        for blk, pk, (res, si), out in zip(self.blocks[gen], self.packs[gen], self.residents[gen], self._outs[gen]):
            st = self.streams[si]
            if pk is not None:
                if pack:
                    pk.pack(blk)
                res.upload_packed(pk, stream=st)
            else:
                res.upload(blk, stream=st)
            res.eval(precision, conv, stream=st)
            res.fetch_async(out, stream=st)

    def wait(self, gen: int = 0):
        """Block until generation `gen` is back in host memory (stream order: everything submitted before it too)."""
        for si in sorted({si for _, si in self.residents[gen]}):
            _check(self._L.pdeb_stream_sync(self.streams[si]), "pdeb_stream_sync")
        return self.robustness[gen], self.status[gen]

    def run(self, precision: int = 64, conv=None, gen: int = 0):
        self.submit(gen, precision, conv)
        return self.wait(gen)

    def _sync_all(self):
        for st in self.streams:
            _check(self._L.pdeb_stream_sync(st), "pdeb_stream_sync")

    @property
    def h2d_bytes(self):
        if self.packed:
            return sum(p.nbytes for p in self.packs[0])
        return sum(b.slab_bytes for b in self.blocks[0])

    def close(self):
        for gen in self.residents:
            for r, _ in gen:
                r.close()
        self.residents = []
        for st in self.streams:
            if st:
                self._L.pdeb_stream_destroy(st)
        self.streams = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def search_batch(batch: ProblemBatch, samples=None, population: int = 4096, generations: int = 6, elite_frac: float = 0.02,
                 bounds=(0.0, 1.0e6, -5.0e3, 5.0e3), seed: int = 0, device: int = 0, conv=None, nthreads: int = 0, knot_y_out=None,
                 timing: bool = False):
    """pdeb_controls_search_batch: the device-resident search for many samples of `batch` in one call (host threads inside
    the library, nothing per sample in Python).  The knot times are the batch's (`n_knots`, `knot_t`; a uniform-knots batch:
    linspace(0, T, n_knots)); the values found are written into `knot_y_out` (default: batch.knot_y itself, so the batch
    can be evaluated right away).  -> dict(best f64[n], found bool[n], kernel_ms, launches); entries of samples that were
    not asked for, or that are outside the controls family, have found False.  kernel_ms is the sum of the searches' device
    times and only measured with timing=True (two more driver calls per search)."""
    n = batch.n
    idx = None if samples is None else np.ascontiguousarray(samples, dtype=np.int64)
    ky = batch.knot_y if knot_y_out is None else knot_y_out
    if ky.dtype != np.float64 or not ky.flags.c_contiguous or ky.shape != batch.knot_y.shape:
        raise ValueError("knot_y_out must look like batch.knot_y")
    best = np.full(n, abi.FAILED_ROBUSTNESS, dtype=np.float64)
    found = np.zeros(n, dtype=np.int32)
    b4 = np.ascontiguousarray(bounds, dtype=np.float64)
    if b4.shape != (4,):
        raise ValueError("bounds = (heat lo, heat hi, wave lo, wave hi)")
    tm = abi.PdebTiming()
    cb = batch.as_c()
    _check(lib().pdeb_controls_search_batch(C.addressof(cb), _ptr(idx), n if idx is None else int(idx.size), int(population),
                                            int(generations), float(elite_frac), _ptr(b4), C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF),
                                            int(device), _conv_ptr(conv), int(nthreads), _ptr(ky), _ptr(best), _ptr(found),
                                            C.addressof(tm) if timing else None), "pdeb_controls_search_batch")
    return {"best": best, "found": found.astype(bool), "kernel_ms": float(tm.kernel_ms), "launches": int(tm.launches)}


class ControlsEvaluator:
    """One specification against many candidate controls (pdeb_controls_*).

    `batch`/`sample` name the specification (its own knots are ignored), `knot_t` the knot times every candidate
    shares, `capacity` the largest number of candidates of one eval() call."""

    def __init__(self, batch: ProblemBatch, knot_t, capacity: int, sample: int = 0, device: int = 0, conv=None):
        self._L = lib()
        self.knot_t = np.ascontiguousarray(knot_t, dtype=np.float64).ravel()
        self.nk = int(self.knot_t.size)
        self.capacity = int(capacity)
        self.n_pred = int(batch.n_pred[sample])
        self._h = C.c_void_p()
        cb = batch.as_c()
        _check(self._L.pdeb_controls_create(C.addressof(cb), sample, self.capacity, self.nk, _ptr(self.knot_t), device,
                                            _conv_ptr(conv), C.addressof(self._h)), "pdeb_controls_create")

    def eval(self, knot_y: Optional[np.ndarray], m: Optional[int] = None, want_robustness: bool = True,
             want_clauses: bool = False, stream: int = 0, out=None):
        """knot_y: [n_knots, m] (plane-major, C-contiguous) or None to re-evaluate the values already on the device.
        out: optional dict of preallocated result arrays ("robustness" f64[m], "status" i32[m]); page-locked ones
        (PinnedBuffer) keep the download at PCIe speed and asynchronous to other streams.
        Returns dict(robustness, status, [clause], best_index, best_value, kernel_ms, launches)."""
        if knot_y is not None:
            knot_y = np.ascontiguousarray(knot_y, dtype=np.float64)
            if knot_y.ndim != 2 or knot_y.shape[0] != self.nk:
                raise ValueError(f"knot_y must be [n_knots={self.nk}, m]")
            m = knot_y.shape[1]
        if m is None or not (1 <= m <= self.capacity):
            raise ValueError("number of candidates out of range")
        if out is not None and want_robustness:
            rob, st = out["robustness"][:m], out["status"][:m]
            if rob.dtype != np.float64 or st.dtype != np.int32 or not rob.flags.c_contiguous or not st.flags.c_contiguous:
                raise ValueError("out arrays must be contiguous f64 / i32")
        else:
            rob = np.empty(m, dtype=np.float64) if want_robustness else None
            st = np.empty(m, dtype=np.int32) if want_robustness else None
        cl = np.empty((self.n_pred, m), dtype=np.float64) if want_clauses else None
        bi, bv, tm = C.c_int64(-1), C.c_double(0.0), abi.PdebTiming()
        _check(self._L.pdeb_controls_eval(self._h, m, _ptr(knot_y), _ptr(rob), _ptr(cl), _ptr(st), C.addressof(bi),
                                          C.addressof(bv), stream or None, C.addressof(tm)), "pdeb_controls_eval")
        res = {"robustness": rob, "status": st, "best_index": int(bi.value), "best_value": float(bv.value),
               "kernel_ms": float(tm.kernel_ms), "launches": int(tm.launches)}
        if want_clauses:
            res["clause"] = cl
        return res

    def respec(self, batch: ProblemBatch, knot_t, sample: int = 0, conv=None):
        """Point the evaluator at another specification of the family (no allocation: what a sweep over many programs
        calls between searches).  Up to 32 knots; the capacity stays."""
        self.knot_t = np.ascontiguousarray(knot_t, dtype=np.float64).ravel()
        self.nk = int(self.knot_t.size)
        self.n_pred = int(batch.n_pred[sample])
        cb = batch.as_c()
        _check(self._L.pdeb_controls_respec(self._h, C.addressof(cb), sample, self.nk, _ptr(self.knot_t), _conv_ptr(conv)),
               "pdeb_controls_respec")

    def search(self, population: int, generations: int, elite_frac: float, lo: float, hi: float, seed: int = 0, stream: int = 0):
        """The whole cross-entropy search on the device (pdeb_controls_search): one launch per generation, nothing
        crosses PCIe until it is over.  -> dict(knot_y [n_knots] or None, robustness, history [generations], mean, std,
        kernel_ms, launches)."""
        nk = self.nk
        best = np.zeros(max(nk, 1), dtype=np.float64)
        mean, std = np.zeros(max(nk, 1), dtype=np.float64), np.zeros(max(nk, 1), dtype=np.float64)
        hist = np.zeros(int(generations), dtype=np.float64)
        bv, found, tm = C.c_double(0.0), C.c_int32(0), abi.PdebTiming()
        _check(self._L.pdeb_controls_search(self._h, int(population), int(generations), float(elite_frac), float(lo), float(hi),
                                            C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), _ptr(best), C.addressof(bv), C.addressof(found),
                                            _ptr(hist), _ptr(mean), _ptr(std), stream or None, C.addressof(tm)), "pdeb_controls_search")
        ok = bool(found.value)
        return {"knot_y": best[:nk].copy() if ok else None, "robustness": float(bv.value) if ok else -np.inf,
                "history": hist.tolist(), "mean": mean[:nk], "std": std[:nk], "kernel_ms": float(tm.kernel_ms),
                "launches": int(tm.launches)}

    def fetch_knots(self, m: int, stream: int = 0) -> np.ndarray:
        """The candidates currently on the device, [n_knots, m]."""
        out = np.zeros((max(self.nk, 1), int(m)), dtype=np.float64)
        _check(self._L.pdeb_controls_fetch_knots(self._h, int(m), _ptr(out), stream or None), "pdeb_controls_fetch_knots")
        return out[: self.nk]

    def close(self):
        if self._h:
            self._L.pdeb_controls_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

