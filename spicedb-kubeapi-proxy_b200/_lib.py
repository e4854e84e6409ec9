"""ctypes binding of libzgpu.so (include/zgpu.h). No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

NO_PERMISSION, HAS_PERMISSION, ITEM_ERROR = 1, 2, 255
SREL_NONE, SREL_WILDCARD = 0xFFFF, 0xFFFE
NO_OBJECT = 0xFFFFFFFF
OP_TOUCH, OP_CREATE, OP_DELETE = 0, 1, 2
PRECOND_MUST_MATCH, PRECOND_MUST_NOT_MATCH = 1, 2
E2BIG = -7

CHECK_DTYPE = np.dtype(
    [("res", "<u4"), ("subj", "<u4"), ("perm", "<u2"), ("stype", "<u2"), ("srel", "<u2"), ("flags", "<u2")]
)
TUPLE_DTYPE = np.dtype(
    [("res", "<u4"), ("subj", "<u4"), ("rel", "<u2"), ("stype", "<u2"), ("srel", "<u2"), ("flags", "<u2")]
)


class ZgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"zgpu error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("subquery_capacity", C.c_uint64),
                ("work_budget", C.c_uint32), ("shard_rank", C.c_uint16), ("shard_count", C.c_uint16),
                ("n_devices", C.c_uint32), ("reserved", C.c_uint32)]


class _RelStr(C.Structure):
    _fields_ = [(n, C.c_char_p) for n in ("res_type", "res_id", "relation", "subj_type", "subj_id", "subj_rel")]


class _UpdateStr(C.Structure):
    _fields_ = [("rel", _RelStr), ("expires_at", C.c_uint32), ("op", C.c_uint32)]


class _Precond(C.Structure):
    _fields_ = [("op", C.c_uint32), ("filter", _RelStr)]


class _ListTemplate(C.Structure):
    _fields_ = [(k, C.c_char_p) for k in ("res_type", "permission", "subj_type", "subj_id", "subj_rel", "req_name",
                                          "req_namespace")] + [("id_kind", C.c_uint32), ("flags", C.c_uint32)]


ID_NAME, ID_NAMESPACED_NAME = 0, 1
TPL_CLEAR_NAMESPACE = 1


class _Stats(C.Structure):
    _fields_ = [("checks", C.c_uint64), ("launches", C.c_uint64), ("passes", C.c_uint64), ("tuples", C.c_uint64),
                ("snapshot_bytes", C.c_uint64), ("revision", C.c_uint64), ("last_alg_bytes", C.c_uint64),
                ("last_kernel_ms", C.c_double), ("coalesced_launches", C.c_uint64), ("coalesced_requests", C.c_uint64),
                ("stack_spills", C.c_uint64), ("memo_batches", C.c_uint64), ("split_batches", C.c_uint64),
                ("delta_publishes", C.c_uint64), ("full_publishes", C.c_uint64), ("last_publish_ms", C.c_double),
                ("streamed_calls", C.c_uint64), ("lookup_batches", C.c_uint64), ("lookups_batched", C.c_uint64),
                ("devices", C.c_uint64)]


class _Update(C.Structure):
    _fields_ = [("res", C.c_uint32), ("subj", C.c_uint32), ("rel", C.c_uint16), ("stype", C.c_uint16),
                ("srel", C.c_uint16), ("flags", C.c_uint16), ("expires_at", C.c_uint32), ("op", C.c_uint32)]


UPDATE_DTYPE = np.dtype([("res", "<u4"), ("subj", "<u4"), ("rel", "<u2"), ("stype", "<u2"), ("srel", "<u2"),
                         ("flags", "<u2"), ("expires_at", "<u4"), ("op", "<u4")])


class _ListItem(C.Structure):
    _fields_ = [("begin", C.c_uint64), ("end", C.c_uint64), ("name_off", C.c_uint64), ("ns_off", C.c_uint64),
                ("name_len", C.c_uint32), ("ns_len", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


LIST_ITEM_DTYPE = np.dtype([("begin", "<u8"), ("end", "<u8"), ("name_off", "<u8"), ("ns_off", "<u8"),
                            ("name_len", "<u4"), ("ns_len", "<u4"), ("flags", "<u4"), ("reserved", "<u4")])
ITEM_IS_OBJECT, ITEM_HAS_METADATA, ITEM_HAS_OBJECT, ITEM_RAW_NAMES = 1, 2, 4, 8
LIST_ITEMS, LIST_TABLE_ROWS, LIST_PROTOBUF, LIST_PROTOBUF_OBJECT = 0, 1, 2, 3
LIST_EMPTY_AS_NULL = 1


def library_path() -> str:
    # ZGPU_LIB selects an alternative build of the same sources (tuning experiments)
    return os.environ.get("ZGPU_LIB") or os.path.join(_HERE, "libzgpu.so")


def build_library(force: bool = False) -> str:
    """nvcc-compile csrc/ into libzgpu.so for sm_100a (cross-compiles without a GPU)."""
    so = library_path()
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cc", ".h", ".cuh"))]
    srcs.append(os.path.join(_ROOT, "include", "zgpu.h"))
    stale = not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if os.environ.get("ZGPU_LIB"):
        stale = force = False  # explicitly chosen library: use as is
    if force or stale:
        if not os.path.exists("/usr/local/cuda/bin/nvcc"):
            if os.path.exists(so):
                return so
            raise ZgpuError(-4, "libzgpu.so is missing and nvcc is not available to build it")
        # several ranks (torchrun) or test workers may import at once: one builds, the others wait
        import fcntl

        with open(os.path.join(csrc, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                still_stale = not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
                if force or still_stale:
                    r = subprocess.run(["make", "-C", csrc], capture_output=True, text=True)
                    if r.returncode != 0:
                        raise ZgpuError(-4, "building libzgpu.so failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return so


_LIB = None

_SIGS = {
    "zg_engine_create": (C.c_int, [C.POINTER(_Config), C.POINTER(C.c_void_p)]),
    "zg_engine_destroy": (None, [C.c_void_p]),
    "zg_last_error": (C.c_char_p, []),
    "zg_build_info": (C.c_char_p, []),
    "zg_last_error_copy": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "zg_clear_relationships": (C.c_int, [C.c_void_p]),
    "zg_load_schema": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "zg_num_types": (C.c_int, [C.c_void_p]),
    "zg_num_slots": (C.c_int, [C.c_void_p]),
    "zg_type_id": (C.c_int, [C.c_void_p, C.c_char_p]),
    "zg_slot_id": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "zg_slot_type": (C.c_int, [C.c_void_p, C.c_int]),
    "zg_slot_is_permission": (C.c_int, [C.c_void_p, C.c_int]),
    "zg_slot_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "zg_type_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "zg_intern_object": (C.c_uint32, [C.c_void_p, C.c_int, C.c_char_p]),
    "zg_find_object": (C.c_uint32, [C.c_void_p, C.c_int, C.c_char_p]),
    "zg_object_name": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_char_p, C.c_size_t]),
    "zg_load_tuples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "zg_apply_updates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "zg_publish": (C.c_int, [C.c_void_p]),
    "zg_num_tuples": (C.c_uint64, [C.c_void_p]),
    "zg_set_clock": (None, [C.c_void_p, C.c_int64]),
    "zg_write_relationships": (C.c_int, [C.c_void_p, C.POINTER(_UpdateStr), C.c_uint64, C.POINTER(_Precond), C.c_uint64]),
    "zg_delete_relationships": (C.c_int, [C.c_void_p, C.POINTER(_RelStr), C.POINTER(_Precond), C.c_uint64,
                                          C.POINTER(C.c_uint64)]),
    "zg_read_relationships": (C.c_int, [C.c_void_p, C.POINTER(_RelStr), C.c_char_p, C.c_size_t,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]),
    "zg_watch_read": (C.c_int, [C.c_void_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t),
                                C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zg_check_bulk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "zg_check_bulk_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "zg_check_bulk_str": (C.c_int, [C.c_void_p, C.POINTER(_RelStr), C.c_uint64, C.c_void_p]),
    "zg_resolve_checks": (C.c_int, [C.c_void_p, C.POINTER(_RelStr), C.c_uint64, C.c_void_p]),
    "zg_resolve_checks_packed": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "zg_check_bulk_packed": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "zg_lookup_resources": (C.c_int, [C.c_void_p, C.c_uint16, C.c_uint16, C.c_uint16, C.c_uint32, C.c_uint16,
                                      C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "zg_lookup_resources_str": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                          C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]),
    "zg_shard_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "zg_shard_subqueries": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]),
    "zg_shard_fold": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]),
    "zg_shard_route_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "zg_shard_pass_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "zg_shard_fold_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]),
    "zg_shard_unroute_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "zg_debug_row": (C.c_int, [C.c_void_p, C.c_uint16, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                               C.POINTER(C.c_uint64)]),
    "zg_host_alloc": (C.c_void_p, [C.c_size_t]),
    "zg_host_free": (None, [C.c_void_p]),
    "zg_list_scan": (C.c_int64, [C.c_char_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint64)]),
    "zg_list_filter": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                                 C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "zg_list_resolve": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint64, C.POINTER(_ListTemplate),
                                  C.c_void_p, C.c_void_p]),
    "zg_list_postfilter": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(_ListTemplate), C.c_uint32,
                                     C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "zg_list_keep_allowed": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_uint32, C.c_char_p,
                                       C.c_char_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_void_p]),
    "zg_list_prefilter": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint32, C.POINTER(_ListTemplate), C.c_void_p,
                                    C.c_size_t, C.POINTER(C.c_size_t)]),
    "zg_stats_get": (C.c_int, [C.c_void_p, C.POINTER(_Stats)]),
    "zg_count_alg_bytes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
}


def exported_symbols():
    """Names include/zgpu.h declares; tests assert the .so exports each of them."""
    return sorted(_SIGS)


def lib():
    """Loads libzgpu.so. There is NO fallback: a missing library is an error."""
    global _LIB
    if _LIB is None:
        so = build_library()
        L = C.CDLL(so)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _b(s):
    if s is None:
        return None
    return s.encode() if isinstance(s, str) else s


def _relstr(rt, rid, rel, st, sid, srel):
    return _RelStr(_b(rt), _b(rid), _b(rel), _b(st), _b(sid), _b(srel or ""))


def split_rel(rel: str):
    """'type:id#rel@stype:sid[#srel]' (pkg/rules/rules.go:1050 grammar) -> 6-tuple."""
    left, right = rel.split("@", 1)
    rt, rest = left.split(":", 1)
    rid, r = rest.rsplit("#", 1)
    st, srest = right.split(":", 1)
    sid, _, srel = srest.partition("#")
    return rt, rid, r, st, sid, srel


def list_scan(body: bytes, mode: int = LIST_ITEMS):
    """zg_list_scan: (items structured array, items_begin, items_end), or None when the body has no
    top-level "items" ("rows" with LIST_TABLE_ROWS) array. LIST_PROTOBUF: a protobuf-encoded <Kind>List
    (items = whole `items` entries, items_begin = offset of raw's length, items_end = end of raw).
    Raises ZgpuError(-1) on a malformed body."""
    L = lib()
    ib, ie = C.c_uint64(0), C.c_uint64(0)
    cap = len(body) // 128 + 16  # an item is rarely shorter; untouched pages of the buffer cost nothing
    while True:
        items = np.empty(cap, dtype=LIST_ITEM_DTYPE)
        n = L.zg_list_scan(body, len(body), mode, items.ctypes.data, cap, C.byref(ib), C.byref(ie))
        if n == -7:
            cap *= 4
            continue
        if n < 0:
            raise ZgpuError(int(n), "malformed list body")
        if ie.value == 0:
            return None
        return items[:n], ib.value, ie.value


def list_filter(body: bytes, items: np.ndarray, keep: np.ndarray, items_begin: int, items_end: int,
                flags: int = 0) -> bytes:
    """zg_list_filter: the body with only the kept items, every other byte untouched."""
    L = lib()
    items = np.ascontiguousarray(items, dtype=LIST_ITEM_DTYPE)
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    out = np.empty(len(body) + 8, dtype=np.uint8)  # the filtered body is never longer than body + 2
    need = C.c_size_t(0)
    rc = L.zg_list_filter(body, len(body), items.ctypes.data, len(items), keep.ctypes.data, items_begin, items_end,
                          flags, out.ctypes.data, out.size, C.byref(need))
    if rc != 0:
        raise ZgpuError(rc, "zg_list_filter failed")
    return out[:need.value].tobytes()


class PinnedArray:
    """numpy view over zg_host_alloc memory (freed with the object)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        self._L = lib()
        self.ptr = self._L.zg_host_alloc(max(n, 1))
        if not self.ptr:
            raise ZgpuError(-8, "zg_host_alloc failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        if getattr(self, "ptr", None):
            self.array = None
            self._L.zg_host_free(self.ptr)
            self.ptr = None


class Engine:
    """One zg_engine: schema + relationship store + published CSR snapshot in HBM."""

    def __init__(self, schema: str | None = None, device: int = -1, subquery_capacity: int = 0, work_budget: int = 0,
                 host_only: bool = False, forward_only: bool = False, shard_rank: int = 0, shard_count: int = 0,
                 n_devices: int = 0):
        """host_only=True builds schema/store/snapshot without a GPU (CPU unit tests of
        the host logic); every check/lookup on such an engine raises ZgpuError."""
        self._L = lib()
        cfg = _Config(device, (1 if host_only else 0) | (2 if forward_only else 0), subquery_capacity, work_budget,
                      shard_rank, shard_count, n_devices & 0xFFFFFFFF, 0)
        h = C.c_void_p()
        rc = self._L.zg_engine_create(C.byref(cfg), C.byref(h))
        if rc:
            raise ZgpuError(rc, self._L.zg_last_error().decode())
        self._h = h
        if schema is not None:
            self.load_schema(schema)

    def close(self):
        if getattr(self, "_h", None):
            self._L.zg_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise ZgpuError(rc, self._L.zg_last_error().decode())

    # -- schema ---------------------------------------------------------------
    def load_schema(self, text: str):
        b = _b(text)
        self._ck(self._L.zg_load_schema(self._h, b, len(b)))

    def type_id(self, name):
        return self._L.zg_type_id(self._h, _b(name))

    def slot_id(self, type_name, rel):
        return self._L.zg_slot_id(self._h, self.type_id(type_name), _b(rel))

    def slot_table(self):
        L, h = self._L, self._h
        return [
            (s, L.zg_type_name(h, L.zg_slot_type(h, s)).decode(), L.zg_slot_name(h, s).decode(),
             bool(L.zg_slot_is_permission(h, s)))
            for s in range(L.zg_num_slots(h))
        ]

    def intern(self, type_name, object_id):
        return self._L.zg_intern_object(self._h, self.type_id(type_name), _b(object_id))

    def find(self, type_name, object_id):
        return self._L.zg_find_object(self._h, self.type_id(type_name), _b(object_id))

    # -- store ----------------------------------------------------------------
    def load_tuples(self, tuples: np.ndarray, expires: np.ndarray | None = None):
        t = np.ascontiguousarray(tuples, dtype=TUPLE_DTYPE)
        ex = None if expires is None else np.ascontiguousarray(expires, dtype=np.uint32)
        self._ck(self._L.zg_load_tuples(self._h, t.ctypes.data, None if ex is None else ex.ctypes.data, t.size))

    def add_bulk(self, type_name, rel, subj_type, res, subj, srel=None, wildcard=False):
        """Bulk TOUCH of relationships sharing (relation, subject type, subject relation)."""
        res = np.ascontiguousarray(res, dtype=np.uint32)
        t = np.zeros(res.size, dtype=TUPLE_DTYPE)
        t["res"] = res
        t["subj"] = 0 if wildcard else np.ascontiguousarray(subj, dtype=np.uint32)
        t["rel"] = self.slot_id(type_name, rel)
        t["stype"] = self.type_id(subj_type)
        t["srel"] = SREL_WILDCARD if wildcard else (SREL_NONE if srel is None else self.slot_id(subj_type, srel))
        self.load_tuples(t)

    def publish(self):
        self._ck(self._L.zg_publish(self._h))

    def apply_updates(self, updates: np.ndarray):
        """zg_apply_updates: interned CREATE / TOUCH / DELETE (UPDATE_DTYPE); visible after publish()."""
        u = np.ascontiguousarray(updates, dtype=UPDATE_DTYPE)
        self._ck(self._L.zg_apply_updates(self._h, u.ctypes.data, u.size))

    def num_tuples(self):
        return self._L.zg_num_tuples(self._h)

    def set_clock(self, unix_seconds: int):
        self._L.zg_set_clock(self._h, int(unix_seconds))

    def write_relationships(self, updates, preconditions=()):
        """updates: [(op, 'type:id#rel@stype:sid[#srel]', expires_at)];
        preconditions: [(op, {filter fields})]"""
        n = len(updates)
        arr = (_UpdateStr * max(n, 1))()
        keep = []
        for i, (op, rel, exp) in enumerate(updates):
            parts = [_b(x) for x in split_rel(rel)]
            keep.append(parts)
            arr[i] = _UpdateStr(_RelStr(*parts), int(exp), int(op))
        pre = (_Precond * max(len(preconditions), 1))()
        for i, (op, f) in enumerate(preconditions):
            pre[i] = _Precond(int(op), self._filter(f))
        self._ck(self._L.zg_write_relationships(self._h, arr, n, pre, len(preconditions)))

    @staticmethod
    def _filter(f):
        return _RelStr(_b(f.get("res_type", "")), _b(f.get("res_id", "")), _b(f.get("rel", "")),
                       _b(f.get("subj_type", "")), _b(f.get("subj_id", "")), _b(f.get("subj_rel", "")))

    def delete_relationships(self, flt: dict, preconditions=()):
        pre = (_Precond * max(len(preconditions), 1))()
        for i, (op, f) in enumerate(preconditions):
            pre[i] = _Precond(int(op), self._filter(f))
        n = C.c_uint64(0)
        f = self._filter(flt)
        self._ck(self._L.zg_delete_relationships(self._h, C.byref(f), pre, len(preconditions), C.byref(n)))
        return n.value

    def read_relationships(self, **flt):
        f = self._filter(flt)
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            need, n = C.c_size_t(0), C.c_uint64(0)
            rc = self._L.zg_read_relationships(self._h, C.byref(f), buf, cap, C.byref(need), C.byref(n))
            if rc == E2BIG:
                cap = need.value + 16
                continue
            self._ck(rc)
            return [l for l in buf.value.decode().split("\n") if l]

    # -- hot path ---------------------------------------------------------------
    def watch_read(self, since_revision: int, res_type: str = ""):
        """-> ([(revision, op_name, relationship text, expires_at)], through_revision)."""
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            need, n, through = C.c_size_t(0), C.c_uint64(0), C.c_uint64(0)
            rc = self._L.zg_watch_read(self._h, since_revision, _b(res_type), buf, cap, C.byref(need), C.byref(n),
                                       C.byref(through))
            if rc == -7:
                cap = need.value
                continue
            self._ck(rc)
            out = []
            for line in buf.value.decode().splitlines():
                parts = line.split(" ")
                out.append((int(parts[0]), parts[1], parts[2], int(parts[3]) if len(parts) > 3 else 0))
            return out, through.value

    def check_bulk(self, items: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """CheckBulkPermissions on interned items (HOST buffers; copies inside the call)."""
        items = np.ascontiguousarray(items, dtype=CHECK_DTYPE)
        if out is None:
            out = np.empty(items.size, dtype=np.uint8)
        self._ck(self._L.zg_check_bulk(self._h, items.ctypes.data, items.size, out.ctypes.data))
        return out

    def check_bulk_ptr(self, items_ptr: int, n: int, out_ptr: int):
        """Raw HOST pointers (e.g. pinned buffers from PinnedArray)."""
        self._ck(self._L.zg_check_bulk(self._h, items_ptr, n, out_ptr))

    def check_bulk_device(self, d_items_ptr: int, n: int, d_out_ptr: int, stream: int = 0):
        """DEVICE pointers (e.g. torch tensors' data_ptr()) on a cudaStream_t handle."""
        self._ck(self._L.zg_check_bulk_device(self._h, d_items_ptr, n, d_out_ptr, stream))

    @staticmethod
    def _relstr_array(rels):
        n = len(rels)
        arr = (_RelStr * max(n, 1))()
        keep = []  # the bytes objects must outlive the call
        for i, r in enumerate(rels):
            parts = [_b(x) for x in (split_rel(r) if isinstance(r, str) else r)]
            keep.append(parts)
            arr[i] = _RelStr(*parts)
        return arr, keep

    def check_bulk_str(self, rels) -> np.ndarray:
        n = len(rels)
        arr, _keep = self._relstr_array(rels)
        out = np.empty(n, dtype=np.uint8)
        self._ck(self._L.zg_check_bulk_str(self._h, arr, n, out.ctypes.data))
        return out

    @staticmethod
    def list_template(res_type, permission, subj_type, subj_id, subj_rel="", id_kind=ID_NAMESPACED_NAME, req_name="",
                      req_namespace="", clear_namespace=False):
        return _ListTemplate(_b(res_type), _b(permission), _b(subj_type), _b(subj_id), _b(subj_rel or ""), _b(req_name),
                             _b(req_namespace), id_kind, TPL_CLEAR_NAMESPACE if clear_namespace else 0)

    def list_resolve(self, body: bytes, items: np.ndarray, tpl: "_ListTemplate"):
        """-> (interned checks, checked mask) for the scanned items; no GPU work."""
        items = np.ascontiguousarray(items, dtype=LIST_ITEM_DTYPE)
        out = np.zeros(len(items), dtype=CHECK_DTYPE)
        checked = np.zeros(len(items), dtype=np.uint8)
        self._ck(self._L.zg_list_resolve(self._h, body, len(body), items.ctypes.data, len(items), C.byref(tpl),
                                         out.ctypes.data, checked.ctypes.data))
        return out, checked

    def list_postfilter(self, body: bytes, templates) -> bytes:
        """zg_list_postfilter: the whole post-filter of a list body in one call."""
        arr = (_ListTemplate * max(len(templates), 1))(*templates)
        out = np.empty(len(body) + 8, dtype=np.uint8)
        need = C.c_size_t(0)
        self._ck(self._L.zg_list_postfilter(self._h, body, len(body), arr, len(templates), out.ctypes.data, out.size,
                                            C.byref(need)))
        return out[:need.value].tobytes()

    def list_keep_allowed(self, body: bytes, items: np.ndarray, res_type: str, allowed_ids: np.ndarray, mode: int = LIST_ITEMS,
                          req_namespace: str = "", self_name: str | None = None) -> np.ndarray:
        """keep mask of the scanned items given the (ascending) object ids a LookupResources returned; no GPU work."""
        items = np.ascontiguousarray(items, dtype=LIST_ITEM_DTYPE)
        allowed = np.ascontiguousarray(allowed_ids, dtype=np.uint32)
        keep = np.zeros(len(items), dtype=np.uint8)
        self._ck(self._L.zg_list_keep_allowed(self._h, body, len(body), items.ctypes.data, len(items), mode, _b(res_type),
                                              _b(req_namespace), allowed.ctypes.data, allowed.size, _b(self_name),
                                              keep.ctypes.data))
        return keep

    def list_prefilter(self, body: bytes, tpl: "_ListTemplate", mode: int = LIST_ITEMS) -> bytes:
        """zg_list_prefilter: LookupResources + scan + keep + splice in one call."""
        out = np.empty(len(body) + 8, dtype=np.uint8)
        need = C.c_size_t(0)
        self._ck(self._L.zg_list_prefilter(self._h, body, len(body), mode, C.byref(tpl), out.ctypes.data, out.size,
                                           C.byref(need)))
        return out[:need.value].tobytes()

    @staticmethod
    def pack_ids(ids):
        """[str] -> (blob, offsets[n+1]) as the packed entry points take them."""
        enc = [_b(i) for i in ids]
        off = np.zeros(len(enc) + 1, dtype=np.uint32)
        np.cumsum([len(x) for x in enc], out=off[1:])
        return b"".join(enc), off

    def _packed_call(self, fn, res_type, relation, subj_type, subj_rel, res_ids, subjects, out):
        blob, off = self.pack_ids(res_ids)
        if isinstance(subjects, (str, bytes)):
            sblob, soff = _b(subjects), None
        else:
            sblob, so = self.pack_ids(subjects)
            soff = so.ctypes.data
            sblob = sblob or b"\0"
        self._ck(fn(self._h, _b(res_type), _b(relation), _b(subj_type), _b(subj_rel or ""), blob or b"\0", off.ctypes.data,
                    sblob, soff, len(res_ids), out.ctypes.data))
        return out

    def resolve_checks_packed(self, res_type, relation, subj_type, res_ids, subjects, subj_rel="") -> np.ndarray:
        """`subjects`: one id (str) for all items, or a list of n ids."""
        return self._packed_call(self._L.zg_resolve_checks_packed, res_type, relation, subj_type, subj_rel, res_ids, subjects,
                                 np.zeros(len(res_ids), dtype=CHECK_DTYPE))

    def check_bulk_packed(self, res_type, relation, subj_type, res_ids, subjects, subj_rel="") -> np.ndarray:
        return self._packed_call(self._L.zg_check_bulk_packed, res_type, relation, subj_type, subj_rel, res_ids, subjects,
                                 np.empty(len(res_ids), dtype=np.uint8))

    def resolve_checks(self, rels) -> np.ndarray:
        """Strings -> interned zg_check items (no GPU work); feed them to check_bulk."""
        n = len(rels)
        arr, _keep = self._relstr_array(rels)
        out = np.zeros(n, dtype=CHECK_DTYPE)
        self._ck(self._L.zg_resolve_checks(self._h, arr, n, out.ctypes.data))
        return out

    def lookup_resources_ids(self, res_type, perm, subj_type, subj, srel=None) -> np.ndarray:
        sr = SREL_NONE if srel is None else self.slot_id(subj_type, srel)
        cap = 1 << 14
        while True:  # on ZG_E2BIG the library keeps the answer: the retry only copies it
            out = np.empty(cap, dtype=np.uint32)
            n = C.c_uint64(0)
            rc = self._L.zg_lookup_resources(self._h, self.type_id(res_type), self.slot_id(res_type, perm),
                                             self.type_id(subj_type), int(subj), sr, out.ctypes.data, cap, C.byref(n))
            if rc == E2BIG:
                cap = int(n.value)
                continue
            self._ck(rc)
            return out[: n.value].copy()

    def lookup_resources_str(self, res_type, perm, subj_type, subj_id, subj_rel=""):
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            need, n = C.c_size_t(0), C.c_uint64(0)
            rc = self._L.zg_lookup_resources_str(self._h, _b(res_type), _b(perm), _b(subj_type), _b(subj_id),
                                                 _b(subj_rel or ""), buf, cap, C.byref(need), C.byref(n))
            if rc == E2BIG:
                cap = need.value + 16
                continue
            self._ck(rc)
            return [l for l in buf.value.decode().split("\n") if l]

    # -- sharded store (dist.ShardedStoreChecker drives these) --------------------
    def shard_pass(self, queries: np.ndarray, level: int) -> int:
        q = np.ascontiguousarray(queries, dtype=CHECK_DTYPE)
        n = C.c_uint64(0)
        self._ck(self._L.zg_shard_pass(self._h, q.ctypes.data, q.size, level, C.byref(n)))
        return n.value

    def shard_subqueries(self, level: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=CHECK_DTYPE)
        self._ck(self._L.zg_shard_subqueries(self._h, level, out.ctypes.data, n))
        return out

    def shard_fold(self, level: int, child_vals: np.ndarray, n_queries: int) -> np.ndarray:
        cv = np.ascontiguousarray(child_vals, dtype=np.uint8)
        out = np.empty(n_queries, dtype=np.uint8)
        self._ck(self._L.zg_shard_fold(self._h, level, cv.ctypes.data, cv.size, out.ctypes.data))
        return out

    # device-resident variants: every pointer is a device address on this engine's GPU (ints, e.g. tensor.data_ptr())
    def shard_route_dev(self, d_items: int, n: int, level: int, n_dest: int, d_routed: int, d_src: int) -> list:
        counts = (C.c_uint64 * n_dest)()
        self._ck(self._L.zg_shard_route_dev(self._h, d_items or None, n, level, n_dest, d_routed or None, d_src or None, counts))
        return [int(c) for c in counts]

    def shard_pass_dev(self, d_queries: int, n: int, level: int) -> int:
        ns = C.c_uint64(0)
        self._ck(self._L.zg_shard_pass_dev(self._h, d_queries or None, n, level, C.byref(ns)))
        return ns.value

    def shard_fold_dev(self, level: int, d_child_vals: int, d_src: int, n_sub: int, d_out: int, final_codes: bool):
        self._ck(self._L.zg_shard_fold_dev(self._h, level, d_child_vals or None, d_src or None, n_sub, d_out or None,
                                           1 if final_codes else 0))

    def shard_unroute_dev(self, d_src: int, d_val: int, n: int, d_out: int):
        self._ck(self._L.zg_shard_unroute_dev(self._h, d_src or None, d_val or None, n, d_out or None))

    def debug_row(self, type_name, rel, res, cls=0, reverse=False) -> np.ndarray:
        """Forward row (resource `res`, class) or, reverse=True, the reverse row of subject `res`."""
        if reverse:
            cls |= 0x80000000
        cap = 1 << 10
        while True:
            out = np.empty(cap, dtype=np.uint32)
            n = C.c_uint64(0)
            rc = self._L.zg_debug_row(self._h, self.slot_id(type_name, rel), int(res), int(cls), out.ctypes.data, cap,
                                      C.byref(n))
            if rc == E2BIG:
                cap = int(n.value)
                continue
            self._ck(rc)
            return out[: n.value].copy()

    # -- measurement ------------------------------------------------------------
    def stats(self) -> dict:
        s = _Stats()
        self._ck(self._L.zg_stats_get(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _Stats._fields_}

    def count_alg_bytes(self, items: np.ndarray) -> int:
        items = np.ascontiguousarray(items, dtype=CHECK_DTYPE)
        b = C.c_uint64(0)
        self._ck(self._L.zg_count_alg_bytes(self._h, items.ctypes.data, items.size, C.byref(b)))
        return b.value
