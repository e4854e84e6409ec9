"""ctypes binding for the CPU oracle (oracle/libzoracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NO_PERMISSION = 1
HAS_PERMISSION = 2
ERROR = 255
SREL_NONE = 0xFFFF
SREL_WILDCARD = 0xFFFE
OP_TOUCH, OP_CREATE, OP_DELETE = 0, 1, 2

# identical to zg_check (include/zgpu.h) and zo_check_item (zanzibar_oracle.h)
CHECK_DTYPE = np.dtype(
    [("res", "<u4"), ("subj", "<u4"), ("perm", "<u2"), ("stype", "<u2"), ("srel", "<u2"), ("flags", "<u2")]
)


def build() -> str:
    """Compile oracle/libzoracle.so if it is missing or stale."""
    so = os.path.join(_HERE, "libzoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("zanzibar_oracle.c", "zanzibar_oracle.h")]
    def stale():
        return not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)

    if stale():
        import fcntl

        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:  # ranks / workers importing at once
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if stale():
                    subprocess.run(["make", "-C", _HERE, "libzoracle.so"], check=True, capture_output=True)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, cp, i, i64, u32, u64 = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64
        sig = {
            "zo_create": (vp, [cp, cp, C.c_size_t]),
            "zo_destroy": (None, [vp]),
            "zo_num_types": (i, [vp]),
            "zo_num_slots": (i, [vp]),
            "zo_type_id": (i, [vp, cp]),
            "zo_slot_id": (i, [vp, i, cp]),
            "zo_slot_type": (i, [vp, i]),
            "zo_slot_is_permission": (i, [vp, i]),
            "zo_slot_name": (cp, [vp, i]),
            "zo_type_name": (cp, [vp, i]),
            "zo_intern": (u32, [vp, i, cp]),
            "zo_find_object": (i64, [vp, i, cp]),
            "zo_object_name": (cp, [vp, i, u32]),
            "zo_write": (i, [vp, i, i, u32, i, u32, i, i64]),
            "zo_write_str": (i, [vp, i, cp, i64]),
            "zo_apply_batch": (i, [vp, vp, u64]),
            "zo_add_bulk": (i, [vp, i, i, i, vp, vp, u64]),
            "zo_num_tuples": (u64, [vp]),
            "zo_check": (i, [vp, vp, i64]),
            "zo_check_str": (i, [vp, cp, cp, cp, cp, cp, cp, i64]),
            "zo_check_bulk": (i, [vp, vp, u64, vp, i, i64]),
            "zo_check_bytes": (u64, [vp, vp, u64, i64]),
            "zo_lookup_resources": (i, [vp, i, i, i, u32, i, i64, vp, u64, C.POINTER(u64)]),
            "zo_read_str": (i64, [vp, cp, cp, cp, cp, cp, cp, i64, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
            "zo_last_error": (cp, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _b(s):
    return s.encode() if isinstance(s, str) else s


class OracleError(RuntimeError):
    pass


class Oracle:
    """Recursive CPU evaluation of Check / LookupResources (see zanzibar_oracle.h)."""

    def __init__(self, schema: str):
        L = _lib()
        err = C.create_string_buffer(512)
        self._h = L.zo_create(_b(schema), err, 512)
        if not self._h:
            raise OracleError(err.value.decode())
        self._L = L

    def close(self):
        if getattr(self, "_h", None):
            self._L.zo_destroy(self._h)
            self._h = None

    __del__ = close

    # -- schema ---------------------------------------------------------
    def type_id(self, name):
        return self._L.zo_type_id(self._h, _b(name))

    def slot_id(self, type_name, rel):
        return self._L.zo_slot_id(self._h, self.type_id(type_name), _b(rel))

    def num_slots(self):
        return self._L.zo_num_slots(self._h)

    def num_types(self):
        return self._L.zo_num_types(self._h)

    def slot_table(self):
        """[(slot, type_name, name, is_permission)] in declaration order."""
        L, h = self._L, self._h
        return [
            (s, L.zo_type_name(h, L.zo_slot_type(h, s)).decode(), L.zo_slot_name(h, s).decode(),
             bool(L.zo_slot_is_permission(h, s)))
            for s in range(L.zo_num_slots(h))
        ]

    def intern(self, type_name, object_id):
        return self._L.zo_intern(self._h, self.type_id(type_name), _b(object_id))

    def object_name(self, type_name, oid):
        r = self._L.zo_object_name(self._h, self.type_id(type_name), oid)
        return r.decode() if r else None

    def _err(self):
        return self._L.zo_last_error(self._h).decode()

    # -- mutations ------------------------------------------------------
    def write(self, rel: str, op=OP_TOUCH, expires_at=0):
        rc = self._L.zo_write_str(self._h, op, _b(rel), int(expires_at))
        if rc:
            raise OracleError(self._err())

    def touch(self, rel, expires_at=0):
        self.write(rel, OP_TOUCH, expires_at)

    def delete(self, rel):
        self.write(rel, OP_DELETE)

    def write_ids(self, op, rel_slot, res, stype, subj, srel=SREL_NONE, expires_at=0):
        """One interned relationship update (same fields as zg_update)."""
        rc = self._L.zo_write(self._h, int(op), int(rel_slot), int(res), int(stype), int(subj), int(srel), int(expires_at))
        if rc:
            raise OracleError(self._err())

    def apply_updates(self, ups: np.ndarray):
        """A batch of interned updates (zg_update records) in one pass over the store."""
        u = np.ascontiguousarray(ups)
        assert u.dtype.itemsize == 24
        if self._L.zo_apply_batch(self._h, u.ctypes.data, u.size):
            raise OracleError(self._err())

    def add_bulk(self, type_name, rel, subj_type, res, subj, srel=None, wildcard=False):
        res = np.ascontiguousarray(res, dtype=np.uint32)
        subj = np.ascontiguousarray(subj, dtype=np.uint32)
        assert res.shape == subj.shape
        sr = SREL_WILDCARD if wildcard else (SREL_NONE if srel is None else self.slot_id(subj_type, srel))
        rc = self._L.zo_add_bulk(self._h, self.slot_id(type_name, rel), self.type_id(subj_type), sr,
                                 res.ctypes.data, subj.ctypes.data, res.size)
        if rc:
            raise OracleError(self._err())

    def num_tuples(self):
        return self._L.zo_num_tuples(self._h)

    # -- queries --------------------------------------------------------
    def check(self, res_type, res_id, perm, subj_type, subj_id, subj_rel="", now=0):
        return self._L.zo_check_str(self._h, _b(res_type), _b(res_id), _b(perm), _b(subj_type), _b(subj_id),
                                    _b(subj_rel or ""), int(now))

    def check_rel(self, rel: str, now=0):
        """rel = 'type:id#perm@stype:sid[#srel]'"""
        left, right = rel.split("@", 1)
        rt, rest = left.split(":", 1)
        rid, perm = rest.rsplit("#", 1)
        st, srest = right.split(":", 1)
        sid, _, srel = srest.partition("#")
        return self.check(rt, rid, perm, st, sid, srel, now)

    def check_bulk(self, items: np.ndarray, nthreads=0, now=0) -> np.ndarray:
        items = np.ascontiguousarray(items, dtype=CHECK_DTYPE)
        out = np.empty(items.size, dtype=np.uint8)
        self._L.zo_check_bulk(self._h, items.ctypes.data, items.size, out.ctypes.data, nthreads, int(now))
        return out

    def check_bytes(self, items: np.ndarray, now=0) -> int:
        items = np.ascontiguousarray(items, dtype=CHECK_DTYPE)
        return self._L.zo_check_bytes(self._h, items.ctypes.data, items.size, int(now))

    def lookup_resources_ids(self, res_type, perm, subj_type, subj, srel=None, now=0) -> np.ndarray:
        sr = SREL_NONE if srel is None else self.slot_id(subj_type, srel)
        cap = 1 << 12
        while True:
            out = np.empty(cap, dtype=np.uint32)
            n = C.c_uint64(0)
            rc = self._L.zo_lookup_resources(self._h, self.type_id(res_type), self.slot_id(res_type, perm),
                                             self.type_id(subj_type), int(subj), sr, int(now), out.ctypes.data,
                                             cap, C.byref(n))
            if rc == 0:
                return out[: n.value].copy()
            if rc == -7:
                cap = int(n.value)
                continue
            raise OracleError(self._err())

    def lookup_resources(self, res_type, perm, subj_type, subj_id, subj_rel="", now=0):
        """String form: list of resource object ids (sorted by internal id)."""
        u = self._L.zo_find_object(self._h, self.type_id(subj_type), _b(subj_id))
        ids = self.lookup_resources_ids(res_type, perm, subj_type, 0xFFFFFFFE if u < 0 else u,
                                        subj_rel or None, now)
        names = [self.object_name(res_type, int(i)) for i in ids]
        if u < 0 and subj_rel and subj_type == res_type and \
                self.check(res_type, subj_id, perm, subj_type, subj_id, subj_rel, now) == 2:
            names.append(subj_id)  # never-written userset subject that is a member of its own permission
        return names

    def read(self, res_type="", res_id="", rel="", subj_type="", subj_id="", subj_rel="", now=0):
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            need = C.c_size_t(0)
            n = self._L.zo_read_str(self._h, _b(res_type), _b(res_id), _b(rel), _b(subj_type), _b(subj_id),
                                    _b(subj_rel), int(now), buf, cap, C.byref(need))
            if n == -7:
                cap = need.value + 16
                continue
            return [l for l in buf.value.decode().split("\n") if l]
