"""ctypes binding of the C ABI in include/iqsb.h (libiqs_b200.so).

This is plumbing for the Python harness (tests, bench.py); it adds no behaviour of its own.
There is no fallback: if the CUDA library is missing, ``load()`` raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# IQS_B200_LIB: an alternative build of the same library (kernel tuning variants, tools/fused_variants.sh)
LIB_PATH = os.environ.get("IQS_B200_LIB") or os.path.join(HERE, "lib", "libiqs_b200.so")

F64, F32 = 0, 1
MEM_DEVICE, MEM_MANAGED = 0, 1
SUM, MAX = 0, 1

_lib = None

c_u64, c_uint, c_int, c_dbl, c_vp = ctypes.c_uint64, ctypes.c_uint, ctypes.c_int, ctypes.c_double, ctypes.c_void_p


class IqsbError(RuntimeError):
    pass


class FGate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("control", ctypes.c_int32), ("target", ctypes.c_int32), ("pad", ctypes.c_int32), ("m", c_dbl * 8)]


class Plan(ctypes.Structure):
    _fields_ = [("active", ctypes.c_int32), ("partner", ctypes.c_int32), ("role", ctypes.c_int32), ("nfix", ctypes.c_int32),
                ("pos", ctypes.c_uint32 * 3), ("val", ctypes.c_uint32 * 3), ("extra0", ctypes.c_uint64), ("extra1", ctypes.c_uint64),
                ("npairs", ctypes.c_uint64), ("link_amps", ctypes.c_uint64)]


def plan_global(kind, rank, nranks, M, pos1, pos2):
    """Host-only: the pairs of a global-qubit gate that `rank` updates (include/iqsb.h, iqsb_plan)."""
    pl = Plan()
    _chk(load().iqsb_plan_global(kind, rank, nranks, M, pos1, pos2, ctypes.byref(pl)))
    return pl


class XPlan(ctypes.Structure):
    _fields_ = [("npartners", ctypes.c_int32), ("split_bit", ctypes.c_int32), ("partner", ctypes.c_int32 * 7), ("split_val", ctypes.c_int32 * 7),
                ("mine", ctypes.c_uint64 * 7), ("theirs", ctypes.c_uint64 * 7), ("amps_per_partner", ctypes.c_uint64), ("link_amps", ctypes.c_uint64)]


def plan_exchange(rank, nranks, M, lpos, gpos):
    """Host-only: what `rank` moves when local positions lpos are exchanged with global positions gpos."""
    k = len(lpos)
    a = (ctypes.c_uint * k)(*lpos)
    b = (ctypes.c_uint * k)(*gpos)
    pl = XPlan()
    _chk(load().iqsb_plan_exchange(rank, nranks, M, k, a, b, ctypes.byref(pl)))
    return pl


class PGate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("control", ctypes.c_int32), ("target", ctypes.c_int32), ("diagonal", ctypes.c_int32)]


def plan_placement(place, M, gates, protect_mask=0, last_use=None, min_evict_bit=0):
    """Host-only: (evict positions, bring positions) of the next qubit exchange.  gates = [(kind, control, target, diagonal)]
    in program order, place[position] = physical bit."""
    n = len(place)
    pl = np.ascontiguousarray(place, dtype=np.uint8)
    arr = (PGate * max(1, len(gates)))()
    for i, (kind, c, t, d) in enumerate(gates):
        arr[i].kind, arr[i].control, arr[i].target, arr[i].diagonal = kind, c, t, int(bool(d))
    lu = None if last_use is None else np.ascontiguousarray(last_use, dtype=np.uint64)
    ev, br, k = (ctypes.c_uint * 3)(), (ctypes.c_uint * 3)(), c_int()
    _chk(load().iqsb_plan_placement(pl.ctypes.data_as(c_vp), n, M, arr, len(gates), ctypes.c_uint64(protect_mask),
                                    None if lu is None else lu.ctypes.data_as(c_vp), min_evict_bit, ev, br, ctypes.byref(k)))
    return [int(ev[i]) for i in range(k.value)], [int(br[i]) for i in range(k.value)]


def _fgates(gates):
    arr = (FGate * len(gates))()
    for i, (kind, c, t, m) in enumerate(gates):
        arr[i].kind, arr[i].control, arr[i].target = kind, c, t
        mm = _m(m)
        for k in range(8):
            arr[i].m[k] = mm[k]
    return arr


def plan_fused(gates, log2_local):
    """Host-only: [(first, last, tile positions)] runs that iqsb_fused executes for `gates`."""
    arr = _fgates(gates)
    n = len(gates)
    run_end = (c_int * max(n, 1))()
    tiles = np.zeros(16 * max(n, 1), dtype=np.uint8)
    nruns = c_int()
    _chk(load().iqsb_plan_fused(arr, n, log2_local, run_end, tiles.ctypes.data_as(c_vp), max(n, 1), ctypes.byref(nruns)))
    out, first = [], 0
    for r in range(nruns.value):
        ns = int(tiles[16 * r])
        out.append((first, int(run_end[r]), tiles[16 * r + 1 : 16 * r + 1 + ns].astype(int).tolist()))
        first = int(run_end[r])
    return out


def plan_fused_order(gates, log2_local, reorder=True):
    """Host-only: the plan iqsb_fused executes -- [(gate indices in execution order, tile positions)] per run.
    With reorder, exact X / CNOT gates may move ahead of gates on other qubits (bit-exact commutation)."""
    arr = _fgates(gates)
    n = len(gates)
    order = (c_int * max(n, 1))()
    run_end = (c_int * max(n, 1))()
    tiles = np.zeros(16 * max(n, 1), dtype=np.uint8)
    nruns = c_int()
    _chk(load().iqsb_plan_fused_order(arr, n, log2_local, int(bool(reorder)), order, run_end, tiles.ctypes.data_as(c_vp), max(n, 1), ctypes.byref(nruns)))
    out, first = [], 0
    for r in range(nruns.value):
        ns = int(tiles[16 * r])
        out.append(([int(order[k]) for k in range(first, int(run_end[r]))], tiles[16 * r + 1 : 16 * r + 1 + ns].astype(int).tolist()))
        first = int(run_end[r])
    return out


class FusedTrace(ctypes.Structure):
    _fields_ = [("gate", ctypes.c_int32), ("run", ctypes.c_int32), ("group", ctypes.c_int32),
                ("cls", ctypes.c_uint8), ("tbit", ctypes.c_uint8), ("ckind", ctypes.c_uint8), ("c", ctypes.c_uint8),
                ("trail", ctypes.c_uint8), ("pad", ctypes.c_uint8 * 3)]


def plan_fused_trace(gates, log2_local, reorder=True):
    """Host-only: (list of dicts, one per gate in execution order; list of register positions per group)."""
    arr = _fgates(gates)
    n = len(gates)
    out = (FusedTrace * max(n, 1))()
    gpos = np.zeros(4 * max(n, 1), dtype=np.uint8)
    ng = c_int()
    _chk(load().iqsb_plan_fused_trace(arr, n, log2_local, int(bool(reorder)), out, gpos.ctypes.data_as(c_vp), ctypes.byref(ng)))
    trace = [dict(gate=out[k].gate, run=out[k].run, group=out[k].group, cls=out[k].cls, tbit=out[k].tbit, ckind=out[k].ckind, c=out[k].c, trail=out[k].trail) for k in range(n)]
    groups = [[int(p) for p in gpos[4 * g : 4 * g + 4] if p != 255] for g in range(ng.value)]
    return trace, groups


def plan_fused_dump(gates, log2_local, reorder=True):
    """Host-only: the raw descriptor blob iqsb_fused hands to its kernel (see iqsb_plan_fused_dump)."""
    arr = _fgates(gates)
    used = ctypes.c_size_t()
    _chk(load().iqsb_plan_fused_dump(arr, len(gates), log2_local, int(bool(reorder)), None, 0, ctypes.byref(used)))
    buf = ctypes.create_string_buffer(max(used.value, 1))
    _chk(load().iqsb_plan_fused_dump(arr, len(gates), log2_local, int(bool(reorder)), buf, used.value, ctypes.byref(used)))
    return bytes(buf.raw[: used.value])


def plan_permute_global_bits(rank, nranks, dst_rank_bit):
    """Host-only: (source rank, destination rank, pairwise, identity) of a rank-bit permutation for `rank`."""
    a = np.ascontiguousarray(dst_rank_bit, dtype=np.uint8)
    src, dst, pw, ident = c_int(), c_int(), c_int(), c_int()
    _chk(load().iqsb_plan_permute_global_bits(rank, nranks, a.ctypes.data_as(c_vp), a.size, ctypes.byref(src), ctypes.byref(dst), ctypes.byref(pw), ctypes.byref(ident)))
    return src.value, dst.value, bool(pw.value), bool(ident.value)


def plan_permute(dst_bit):
    """Host-only: list of (positions, dstslot) tile phases for a local qubit permutation."""
    a = np.ascontiguousarray(dst_bit, dtype=np.uint8)
    out = np.zeros(64 * 25, dtype=np.uint8)
    n = c_int()
    _chk(load().iqsb_plan_permute(a.ctypes.data_as(c_vp), a.size, out.ctypes.data_as(c_vp), 64, ctypes.byref(n)))
    phases = []
    for p in range(n.value):
        ns = int(out[25 * p])
        phases.append((out[25 * p + 1 : 25 * p + 1 + ns].astype(int).tolist(), out[25 * p + 13 : 25 * p + 13 + ns].astype(int).tolist()))
    return phases


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IqsbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    L.iqsb_last_error.restype = ctypes.c_char_p
    L.iqsb_launch_count.restype = c_u64
    L.iqsb_nvlink_bytes.restype = c_u64
    L.iqsb_local_amps.restype = c_u64
    L.iqsb_device_ptr.restype = c_vp
    L.iqsb_host_ptr.restype = c_vp
    L.iqsb_get_stream.restype = c_vp
    sig = {
        "iqsb_unique_id": [c_vp],
        "iqsb_init": [c_int, c_int, c_vp, c_int, ctypes.POINTER(c_vp)],
        "iqsb_finalize": [c_vp],
        "iqsb_rank": [c_vp], "iqsb_nranks": [c_vp], "iqsb_device": [c_vp],
        "iqsb_sync": [c_vp],
        "iqsb_mem_info": [c_vp, ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)],
        "iqsb_set_stream": [c_vp, c_vp], "iqsb_get_stream": [c_vp], "iqsb_set_arith": [c_vp, c_int], "iqsb_get_arith": [c_vp],
        "iqsb_launch_count": [c_vp], "iqsb_nvlink_bytes": [c_vp],
        "iqsb_prob_all": [c_vp, c_vp, c_int], "iqsb_pauli_expect": [c_vp, c_u64, c_u64, c_u64, c_u64, c_vp],
        "iqsb_check": [c_vp],
        "iqsb_profile": [c_vp, c_int], "iqsb_profile_read": [c_vp, c_vp, ctypes.c_size_t],
        "iqsb_timer_start": [c_vp], "iqsb_timer_stop": [c_vp, ctypes.POINTER(c_dbl)],
        "iqsb_event_record": [c_vp, c_int], "iqsb_event_elapsed": [c_vp, c_int, c_int, ctypes.POINTER(c_dbl)],
        "iqsb_allreduce_f64": [c_vp, c_vp, c_int, c_int],
        "iqsb_bcast_f64": [c_vp, c_vp, c_int, c_int],
        "iqsb_barrier": [c_vp],
        "iqsb_alloc": [c_vp, c_u64, c_u64, c_int, c_int, ctypes.POINTER(c_vp)],
        "iqsb_free": [c_vp],
        "iqsb_local_amps": [c_vp], "iqsb_dtype": [c_vp], "iqsb_device_ptr": [c_vp], "iqsb_host_ptr": [c_vp],
        "iqsb_prefetch_device": [c_vp],
        "iqsb_upload": [c_vp, c_vp, c_u64, c_u64],
        "iqsb_download": [c_vp, c_vp, c_u64, c_u64],
        "iqsb_copy": [c_vp, c_vp],
        "iqsb_fill_const": [c_vp, c_dbl, c_dbl],
        "iqsb_set_amp": [c_vp, c_u64, c_dbl, c_dbl],
        "iqsb_get_amp": [c_vp, c_u64, ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)],
        "iqsb_fill_random": [c_vp, c_u64, c_u64],
        "iqsb_gate1": [c_vp, c_uint, c_vp, c_u64, c_u64],
        "iqsb_cgate1": [c_vp, c_uint, c_uint, c_vp, c_u64, c_u64],
        "iqsb_swap2x2": [c_vp, c_uint, c_uint, c_vp],
        "iqsb_diag2": [c_vp, c_uint, c_uint, c_vp, c_u64],
        "iqsb_scale": [c_vp, c_vp, c_u64, c_u64],
        "iqsb_phase_by_bit": [c_vp, c_int, c_uint, c_vp, c_vp],
        "iqsb_gate2": [c_vp, c_uint, c_uint, c_vp],
        "iqsb_fused": [c_vp, c_vp, c_int],
        "iqsb_fused_max_log2tile": [c_vp],
        "iqsb_plan_fused": [c_vp, c_int, c_uint, c_vp, c_vp, c_int, ctypes.POINTER(c_int)],
        "iqsb_plan_fused_order": [c_vp, c_int, c_uint, c_int, c_vp, c_vp, c_vp, c_int, ctypes.POINTER(c_int)],
        "iqsb_plan_permute_global_bits": [c_int, c_int, c_vp, c_uint, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)],
        "iqsb_plan_fused_dump": [c_vp, c_int, c_uint, c_int, c_vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)],
        "iqsb_plan_fused_trace": [c_vp, c_int, c_uint, c_int, c_vp, c_vp, ctypes.POINTER(c_int)],
        "iqsb_prob1": [c_vp, c_uint, ctypes.POINTER(c_dbl)],
        "iqsb_parity_expect": [c_vp, c_u64, c_u64, ctypes.POINTER(c_dbl)],
        "iqsb_norm2": [c_vp, ctypes.POINTER(c_dbl)],
        "iqsb_overlap": [c_vp, c_vp, c_vp],
        "iqsb_maxabsdiff": [c_vp, c_vp, c_vp, ctypes.POINTER(c_dbl)],
        "iqsb_l2diff": [c_vp, c_vp, ctypes.POINTER(c_dbl)],
        "iqsb_any_above": [c_vp, c_uint, c_dbl, c_u64, c_vp],
        "iqsb_equal": [c_vp, c_vp, ctypes.POINTER(c_int)],
        "iqsb_entropy_stats": [c_vp, c_vp],
        "iqsb_collapse": [c_vp, c_uint, c_int],
        "iqsb_axpy": [c_vp, c_vp, c_vp],
        "iqsb_qaoa_maxcut": [c_vp, c_uint, c_vp, c_int, c_vp, c_u64, ctypes.POINTER(c_dbl)],
        "iqsb_qaoa_layer": [c_vp, c_vp, c_dbl],
        "iqsb_qaoa_expect": [c_vp, c_vp, c_vp],
        "iqsb_qaoa_histogram": [c_vp, c_vp, c_int, c_dbl, c_dbl, c_vp],
        "iqsb_permute_local": [c_vp, c_vp, c_uint],
        "iqsb_plan_permute": [c_vp, c_uint, c_vp, c_int, ctypes.POINTER(c_int)],
        "iqsb_plan_global": [c_int, c_int, c_int, c_uint, c_uint, c_uint, c_vp],
        "iqsb_share": [c_vp],
        "iqsb_idle_global": [c_vp],
        "iqsb_gate1_global": [c_vp, c_uint, c_uint, c_vp],
        "iqsb_cgate1_global": [c_vp, c_uint, c_uint, c_uint, c_vp],
        "iqsb_swap2x2_global": [c_vp, c_uint, c_uint, c_uint, c_vp],
        "iqsb_permute_global": [c_vp, c_int, c_int], "iqsb_permute_global_bits": [c_vp, c_vp, c_uint],
        "iqsb_exchange_bits": [c_vp, c_uint, c_int, c_vp, c_vp],
        "iqsb_plan_exchange": [c_int, c_int, c_uint, c_int, c_vp, c_vp, c_vp],
        "iqsb_plan_placement": [c_vp, c_uint, c_uint, c_vp, c_int, c_u64, c_vp, c_uint, c_vp, c_vp, ctypes.POINTER(c_int)],
    }
    for name, args in sig.items():
        getattr(L, name).argtypes = args
    _lib = L
    return L


DECLARED_SYMBOLS = None  # filled lazily from include/iqsb.h by tests


def _chk(rc):
    if rc != 0:
        raise IqsbError(f"[{rc}] {load().iqsb_last_error().decode(errors='replace')}")


def _m(m, n=8):
    m = np.asarray(m)
    if m.dtype.kind == "c":
        m = np.ascontiguousarray(m, dtype=np.complex128).ravel().view(np.float64)
    m = np.ascontiguousarray(m, dtype=np.float64).ravel()
    assert m.size == n, (m.size, n)
    return m


def _c2(z):
    z = complex(z)
    return np.array([z.real, z.imag], dtype=np.float64)


def version():
    return int(load().iqsb_version())


def unique_id():
    buf = ctypes.create_string_buffer(128)
    _chk(load().iqsb_unique_id(buf))
    return buf.raw


class Context:
    def __init__(self, rank=0, nranks=1, uid=None, device=-1):
        self.L = load()
        h = c_vp()
        ub = ctypes.create_string_buffer(uid, 128) if uid is not None else None
        _chk(self.L.iqsb_init(rank, nranks, ub, device, ctypes.byref(h)))
        self.h = h
        self.rank, self.nranks = rank, nranks

    def close(self):
        if self.h:
            self.L.iqsb_finalize(self.h)
            self.h = None

    def sync(self):
        _chk(self.L.iqsb_sync(self.h))

    def mem_info(self):
        f, t = c_u64(), c_u64()
        _chk(self.L.iqsb_mem_info(self.h, ctypes.byref(f), ctypes.byref(t)))
        return int(f.value), int(t.value)

    def set_stream(self, cuda_stream_ptr):
        _chk(self.L.iqsb_set_stream(self.h, c_vp(cuda_stream_ptr)))

    def set_arith(self, fma):
        """False: exact, reference operation order (default); True: contracted multiply-adds in the fused kernel"""
        _chk(self.L.iqsb_set_arith(self.h, 1 if fma else 0))

    def get_arith(self):
        return int(self.L.iqsb_get_arith(self.h))

    def launches(self):
        return int(self.L.iqsb_launch_count(self.h))

    def nvlink_bytes(self):
        return int(self.L.iqsb_nvlink_bytes(self.h))

    def check(self):
        """raises IqsbError if a kernel reported a failure (peer barrier deadline)"""
        _chk(self.L.iqsb_check(self.h))

    def profile(self, on):
        """start (clearing) / stop the per-kernel-class device timing"""
        _chk(self.L.iqsb_profile(self.h, int(bool(on))))

    def profile_read(self):
        """{class name: {"launches", "ms", "bytes"}} of the last profiled region"""
        import json

        buf = ctypes.create_string_buffer(1 << 16)
        _chk(self.L.iqsb_profile_read(self.h, buf, len(buf)))
        d = json.loads(buf.value.decode())
        out = {c["name"]: {"launches": c["launches"], "ms": c["ms"], "bytes": c["bytes"]} for c in d["classes"]}
        out["_overflow"] = d["overflow"]
        return out

    def timer_start(self):
        _chk(self.L.iqsb_timer_start(self.h))

    def timer_stop(self):
        ms = c_dbl()
        _chk(self.L.iqsb_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def event_record(self, slot):
        _chk(self.L.iqsb_event_record(self.h, slot))

    def event_elapsed(self, a, b):
        ms = c_dbl()
        _chk(self.L.iqsb_event_elapsed(self.h, a, b, ctypes.byref(ms)))
        return ms.value

    def allreduce(self, values, op=SUM):
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        _chk(self.L.iqsb_allreduce_f64(self.h, v.ctypes.data_as(c_vp), v.size, op))
        return v

    def bcast(self, values, root=0):
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        _chk(self.L.iqsb_bcast_f64(self.h, v.ctypes.data_as(c_vp), v.size, root))
        return v

    def barrier(self):
        _chk(self.L.iqsb_barrier(self.h))

    def alloc(self, local_amps, tmp_amps=0, dtype=F64, mem=MEM_DEVICE):
        return State(self, local_amps, tmp_amps, dtype, mem)


class State:
    """One register shard in HBM.  Method names follow include/iqsb.h."""

    def __init__(self, ctx, local_amps, tmp_amps=0, dtype=F64, mem=MEM_DEVICE):
        self.ctx, self.L = ctx, ctx.L
        h = c_vp()
        _chk(self.L.iqsb_alloc(ctx.h, local_amps, tmp_amps, dtype, mem, ctypes.byref(h)))
        self.h = h
        self.local_amps = int(local_amps)
        self.dtype = dtype
        self.np_dtype = np.complex128 if dtype == F64 else np.complex64

    def free(self):
        if self.h:
            self.L.iqsb_free(self.h)
            self.h = None

    # memory
    def upload(self, host, first=0):
        a = np.ascontiguousarray(host, dtype=self.np_dtype)
        _chk(self.L.iqsb_upload(self.h, a.ctypes.data_as(c_vp), first, a.size))

    def download(self, first=0, count=None):
        count = self.local_amps - first if count is None else count
        out = np.empty(count, dtype=self.np_dtype)
        _chk(self.L.iqsb_download(self.h, out.ctypes.data_as(c_vp), first, count))
        return out

    def copy_from(self, other):
        _chk(self.L.iqsb_copy(self.h, other.h))

    def fill_const(self, z):
        z = complex(z)
        _chk(self.L.iqsb_fill_const(self.h, z.real, z.imag))

    def fill_random(self, seed, global_offset=0):
        _chk(self.L.iqsb_fill_random(self.h, seed, global_offset))

    def set_amp(self, i, z):
        z = complex(z)
        _chk(self.L.iqsb_set_amp(self.h, i, z.real, z.imag))

    def get_amp(self, i):
        re, im = c_dbl(), c_dbl()
        _chk(self.L.iqsb_get_amp(self.h, i, ctypes.byref(re), ctypes.byref(im)))
        return complex(re.value, im.value)

    # gates
    def gate1(self, pos, m, sind=0, eind=None):
        mm = _m(m)
        _chk(self.L.iqsb_gate1(self.h, pos, mm.ctypes.data_as(c_vp), sind, self.local_amps if eind is None else eind))

    def cgate1(self, cpos, tpos, m, sind=0, eind=None):
        mm = _m(m)
        _chk(self.L.iqsb_cgate1(self.h, cpos, tpos, mm.ctypes.data_as(c_vp), sind, self.local_amps if eind is None else eind))

    def swap2x2(self, pos1, pos2, m):
        mm = _m(m)
        _chk(self.L.iqsb_swap2x2(self.h, pos1, pos2, mm.ctypes.data_as(c_vp)))

    def diag2(self, pos1, pos2, d, glb_start=0):
        dd = _m(d)
        _chk(self.L.iqsb_diag2(self.h, pos1, pos2, dd.ctypes.data_as(c_vp), glb_start))

    def scale(self, f, start=0, end=None):
        ff = _c2(f)
        _chk(self.L.iqsb_scale(self.h, ff.ctypes.data_as(c_vp), start, self.local_amps if end is None else end))

    def phase_by_bit(self, cpos, pos, d0, d1):
        a, b = _c2(d0), _c2(d1)
        _chk(self.L.iqsb_phase_by_bit(self.h, cpos, pos, a.ctypes.data_as(c_vp), b.ctypes.data_as(c_vp)))

    def gate2(self, pos_high, pos_low, m16):
        mm = np.ascontiguousarray(np.asarray(m16, dtype=np.complex128).reshape(16)).view(np.float64)
        _chk(self.L.iqsb_gate2(self.h, pos_high, pos_low, mm.ctypes.data_as(c_vp)))

    def prob_all(self):
        """[sum |a|^2, P(bit 0 = 1), P(bit 1 = 1), ...] over the local shard, one read of the state"""
        n = int(self.local_amps).bit_length() - 1
        out = (ctypes.c_double * (n + 1))()
        _chk(self.L.iqsb_prob_all(self.h, out, n + 1))
        return np.array(out[:])

    def pauli_expect(self, xmask, ymask, zmask, glb_start=0, with_norm=False):
        """<psi| X^xmask Y^ymask Z^zmask |psi>, read-only (with_norm: also sum |a|^2 from the same read)"""
        out = (ctypes.c_double * 2)()
        _chk(self.L.iqsb_pauli_expect(self.h, c_u64(xmask), c_u64(ymask), c_u64(zmask), c_u64(glb_start), out))
        return (out[0], out[1]) if with_norm else out[0]

    def fused(self, gates):
        """gates: list of (kind, control, target, m8)."""
        _chk(self.L.iqsb_fused(self.h, _fgates(gates), len(gates)))

    def fused_max_log2tile(self):
        return int(self.L.iqsb_fused_max_log2tile(self.h))

    # reductions
    def prob1(self, pos):
        o = c_dbl()
        _chk(self.L.iqsb_prob1(self.h, pos, ctypes.byref(o)))
        return o.value

    def parity_expect(self, mask, glb_start=0):
        o = c_dbl()
        _chk(self.L.iqsb_parity_expect(self.h, mask, glb_start, ctypes.byref(o)))
        return o.value

    def norm2(self):
        o = c_dbl()
        _chk(self.L.iqsb_norm2(self.h, ctypes.byref(o)))
        return o.value

    def overlap(self, other):
        o = np.zeros(2)
        _chk(self.L.iqsb_overlap(self.h, other.h, o.ctypes.data_as(c_vp)))
        return complex(o[0], o[1])

    def maxabsdiff(self, other, f=1.0):
        o = c_dbl()
        ff = _c2(f)
        _chk(self.L.iqsb_maxabsdiff(self.h, other.h, ff.ctypes.data_as(c_vp), ctypes.byref(o)))
        return o.value

    def l2diff(self, other):
        o = c_dbl()
        _chk(self.L.iqsb_l2diff(self.h, other.h, ctypes.byref(o)))
        return o.value

    def any_above(self, pos, tol, glb_start=0):
        o = (c_int * 2)()
        _chk(self.L.iqsb_any_above(self.h, pos, tol, glb_start, o))
        return int(o[0]), int(o[1])

    def equal(self, other):
        o = c_int()
        _chk(self.L.iqsb_equal(self.h, other.h, ctypes.byref(o)))
        return bool(o.value)

    def entropy_stats(self):
        o = np.zeros(11)
        _chk(self.L.iqsb_entropy_stats(self.h, o.ctypes.data_as(c_vp)))
        return o

    def collapse(self, pos, value):
        _chk(self.L.iqsb_collapse(self.h, pos, int(bool(value))))

    def axpy(self, other, f=1.0):
        ff = _c2(f)
        _chk(self.L.iqsb_axpy(self.h, other.h, ff.ctypes.data_as(c_vp)))

    # QAOA helpers (this shard is the cost-function vector `diag` or the state `psi`)
    def qaoa_maxcut(self, adjacency, weighted=False, pos_of_qubit=None, glb_start=0):
        a = np.ascontiguousarray(adjacency, dtype=np.float64)
        n = a.shape[0]
        pos = np.ascontiguousarray(range(n) if pos_of_qubit is None else pos_of_qubit, dtype=np.uint8)
        o = c_dbl()
        _chk(self.L.iqsb_qaoa_maxcut(self.h, n, a.ctypes.data_as(c_vp), int(bool(weighted)), pos.ctypes.data_as(c_vp), glb_start, ctypes.byref(o)))
        return o.value

    def qaoa_layer(self, diag, gamma):
        _chk(self.L.iqsb_qaoa_layer(self.h, diag.h, gamma))

    def qaoa_expect(self, diag):
        o = np.zeros(2)
        _chk(self.L.iqsb_qaoa_expect(self.h, diag.h, o.ctypes.data_as(c_vp)))
        return float(o[0]), float(o[1])

    def qaoa_histogram(self, diag, nbins, bin_width=1.0, eps=0.0):
        o = np.zeros(nbins)
        _chk(self.L.iqsb_qaoa_histogram(self.h, diag.h, nbins, bin_width, eps, o.ctypes.data_as(c_vp)))
        return o

    def permute_local(self, dst_bit):
        a = np.ascontiguousarray(dst_bit, dtype=np.uint8)
        _chk(self.L.iqsb_permute_local(self.h, a.ctypes.data_as(c_vp), a.size))

    # distributed
    def share(self):
        _chk(self.L.iqsb_share(self.h))

    def gate1_global(self, M, pos, m):
        mm = _m(m)
        _chk(self.L.iqsb_gate1_global(self.h, M, pos, mm.ctypes.data_as(c_vp)))

    def idle_global(self):
        _chk(self.L.iqsb_idle_global(self.h))

    def cgate1_global(self, M, cpos, tpos, m):
        mm = _m(m)
        _chk(self.L.iqsb_cgate1_global(self.h, M, cpos, tpos, mm.ctypes.data_as(c_vp)))

    def swap2x2_global(self, M, pos1, pos2, m):
        mm = _m(m)
        _chk(self.L.iqsb_swap2x2_global(self.h, M, pos1, pos2, mm.ctypes.data_as(c_vp)))

    def exchange_bits(self, M, lpos, gpos):
        k = len(lpos)
        _chk(self.L.iqsb_exchange_bits(self.h, M, k, (ctypes.c_uint * k)(*lpos), (ctypes.c_uint * k)(*gpos)))

    def permute_global(self, src_rank, dst_rank):
        _chk(self.L.iqsb_permute_global(self.h, src_rank, dst_rank))

    def permute_global_bits(self, dst_rank_bit):
        a = np.ascontiguousarray(dst_rank_bit, dtype=np.uint8)
        _chk(self.L.iqsb_permute_global_bits(self.h, a.ctypes.data_as(c_vp), a.size))
