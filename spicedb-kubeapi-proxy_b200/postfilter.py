"""filter_list_response -- host mirror of the reference's list post-filter.

Reference: pkg/authz/postfilter.go:17-55 (filterListResponse) and :58-178
(filterItemsWithBulkPermissions). Same decisions, different mechanics:

  reference                                   here
  json.Unmarshal(body) into maps              zg_list_scan: one structural pass, byte ranges only
  metadata["name"], ["namespace"] per item    offsets recorded by the scan (escapes decoded lazily)
  one CheckBulkPermissions for all items      the same single call (-> one GPU launch)
  json.Marshal(filtered map)                  zg_list_filter: splice of the kept items' bytes

Decisions that are mirrored exactly (each is what the Go code does, not what one might prefer):
  * no top-level "items" array, or an empty one  -> body returned unchanged      (postfilter.go:25-35)
  * an item that is not an object                -> never checked, always kept   (:68-71, :141-146)
  * a template that fails to resolve for an item -> that check is skipped        (:91-95)
  * an item with no checks at all                -> kept                         (:141-146)
  * several post-filters                         -> ALL must be HAS_PERMISSION   (:149-170)
  * a per-pair error                             -> the item is dropped          (:158-162)
  * the bulk call itself failing                 -> the whole filter fails       (:127-130)
  * nothing kept                                 -> "items": null                (:138, nil slice)
  * name/namespace fall back to the request's; `namespaces` lists clear the namespace;
    namespacedName = "ns/name" or "name"                       (pkg/rules/rules.go:312-339)

Only the plain `{{dotted.name}}` form of the reference's relationship templates is resolved here
(name, namespace, namespacedName, resourceId, user.name, user.uid, request.*): the expression
language behind `{{ }}` belongs to the rules engine, which SURVEY.md 8 puts out of scope. A
callable `(fields) -> "type:id#perm@stype:sid"` can be passed instead of a template string.
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Sequence, Union

import numpy as np

from . import _lib
from .client import (LOOKUP_PERMISSIONSHIP_HAS_PERMISSION, PERMISSIONSHIP_HAS_PERMISSION, CheckBulkPermissionsRequest,
                     CheckBulkPermissionsRequestItem, LookupResourcesRequest, ObjectReference, SubjectReference)

_TPL = re.compile(r"\{\{\s*([A-Za-z_][A-Za-z0-9_.]*)\s*\}\}")


@dataclass
class RequestInfo:
    """The fields of k8s.io/apiserver request.RequestInfo the normalisation reads."""
    verb: str = "list"
    resource: str = ""
    name: str = ""
    namespace: str = ""
    api_group: str = ""
    api_version: str = ""


@dataclass
class UserInfo:
    name: str = ""
    uid: str = ""
    groups: List[str] = field(default_factory=list)


class ResolveError(ValueError):
    pass


PostFilter = Union[str, Callable[[Dict[str, str]], str]]


def _fields(req: RequestInfo, user: UserInfo, has_meta: bool, name: str, namespace: str) -> Dict[str, str]:
    # pkg/rules/rules.go:312-339 NewResolveInput
    if not has_meta:
        name = namespace = ""
    name = name or req.name
    namespace = namespace or req.namespace
    if req.resource == "namespaces":
        namespace = ""
    nn = f"{namespace}/{name}" if namespace else name
    return {"name": name, "namespace": namespace, "namespacedName": nn, "resourceId": nn, "user.name": user.name,
            "user.uid": user.uid, "request.verb": req.verb, "request.resource": req.resource,
            "request.name": req.name, "request.namespace": req.namespace, "request.apiGroup": req.api_group,
            "request.apiVersion": req.api_version}


def resolve_rel(f: PostFilter, fields: Dict[str, str]):
    """-> (res_type, res_id, permission, subj_type, subj_id, subj_rel); ResolveError if it cannot."""
    if callable(f):
        rel = f(fields)
    else:
        def sub(m):
            if m.group(1) not in fields:
                raise ResolveError(f"unknown template field {m.group(1)!r}")
            return fields[m.group(1)]
        rel = _TPL.sub(sub, f)
    try:
        parts = _lib.split_rel(rel)
    except ValueError:
        raise ResolveError(f"not a relationship: {rel!r}") from None
    if not all(parts[:5]):
        raise ResolveError(f"relationship with an empty field: {rel!r}")
    return parts


def _text(body: bytes, off: int, ln: int) -> str:
    raw = body[off:off + ln]
    if b"\\" in raw:
        return json.loads(b'"' + raw + b'"')
    return raw.decode("utf-8", "replace")


def filter_list_response(body: bytes, post_filters: Sequence[PostFilter], request: RequestInfo, user: UserInfo,
                         permissions_client) -> bytes:
    """Returns the (possibly filtered) list body. Raises ValueError on a malformed body and whatever
    `permissions_client.CheckBulkPermissions` raises."""
    try:
        scanned = _lib.list_scan(body)
    except _lib.ZgpuError:
        raise ValueError("failed to parse list response") from None
    if scanned is None:
        return body
    items, ib, ie = scanned
    if len(items) == 0:
        return body

    bulk: List[CheckBulkPermissionsRequestItem] = []
    owner: List[int] = []
    for i, it in enumerate(items):
        flags = int(it["flags"])
        if not flags & _lib.ITEM_IS_OBJECT:
            continue
        fields = _fields(request, user, bool(flags & _lib.ITEM_HAS_METADATA),
                         _text(body, int(it["name_off"]), int(it["name_len"])),
                         _text(body, int(it["ns_off"]), int(it["ns_len"])))
        for f in post_filters:
            try:
                rt, rid, perm, st, sid, srel = resolve_rel(f, fields)
            except ResolveError:
                continue
            bulk.append(CheckBulkPermissionsRequestItem(ObjectReference(rt, rid), perm,
                                                        SubjectReference(ObjectReference(st, sid), srel)))
            owner.append(i)

    keep = np.ones(len(items), dtype=np.uint8)
    if bulk:
        resp = permissions_client.CheckBulkPermissions(CheckBulkPermissionsRequest(bulk))
        for k, i in enumerate(owner):
            if k >= len(resp.pairs):
                keep[i] = 0
                continue
            pair = resp.pairs[k]
            item = pair.GetItem()
            if pair.GetError() is not None or item is None or item.permissionship != PERMISSIONSHIP_HAS_PERMISSION:
                keep[i] = 0
    # the reference re-marshals even when everything is kept; the splice then reproduces the body
    return _lib.list_filter(body, items, keep, ib, ie, _lib.LIST_EMPTY_AS_NULL)


# ---- the pre-filter side: LookupResources -> allowed set -> list / table / object ---------------------
# Reference: pkg/authz/lookups.go:19-36 (prefilterResult), :44-132 (runLookupResources),
# pkg/authz/responsefilterer.go:349-415 (filterTable / filterList / filterObject).


def split_name(resource_id: str) -> str:
    """The usual fromObjectIDNameExpr: the part after the last '/' of "namespace/name"."""
    return resource_id.rsplit("/", 1)[-1]


def split_namespace(resource_id: str):
    """The usual fromObjectIDNamespaceExpr: the part before the '/', None for cluster-scoped ids."""
    return resource_id.rsplit("/", 1)[0] if "/" in resource_id else None


class Unauthorized(Exception):
    pass


@dataclass
class PrefilterResult:
    """lookups.go:19-36."""
    all_allowed: bool = False
    allowed_results: set = field(default_factory=set)  # of (namespace, name)

    def IsAllowed(self, namespace: str, name: str) -> bool:
        return self.all_allowed or (namespace, name) in self.allowed_results


def run_lookup_resources(permissions_client, rel, request: RequestInfo,
                         name_from_object_id: Callable[[str], str] = split_name,
                         namespace_from_object_id: Callable[[str], object] = split_namespace) -> PrefilterResult:
    """lookups.go:44-132. `rel` = (res_type, "$", permission, subj_type, subj_id, subj_rel); the two
    callables stand in for the rule's Bloblang expressions (rules engine: out of scope)."""
    rt, rid, perm, st, sid, srel = rel
    if rid != "$":
        raise ValueError("preFilter called with non-$ resource ID")  # lookups.go:45-48
    res = PrefilterResult()
    stream = permissions_client.LookupResources(
        LookupResourcesRequest(rt, perm, SubjectReference(ObjectReference(st, sid), srel or "")))
    for resp in stream:
        if resp.permissionship != LOOKUP_PERMISSIONSHIP_HAS_PERMISSION:
            continue  # conditional results are skipped (lookups.go:86-89)
        name = name_from_object_id(resp.resource_object_id)
        if not name:
            raise ValueError("unable to determine name for resource")  # lookups.go:106-109
        ns = namespace_from_object_id(resp.resource_object_id)
        if ns is None:
            ns = request.namespace or ""  # the expression is re-run on the request input (lookups.go:117-127)
        res.allowed_results.add((ns, name))
    return res


def _filter_by_result(body: bytes, result: PrefilterResult, mode: int, what: str) -> bytes:
    try:
        scanned = _lib.list_scan(body, mode)
    except _lib.ZgpuError:
        raise ValueError(f"failed to decode response body as a {what}") from None
    if scanned is None:
        return body
    items, ib, ie = scanned
    keep = np.zeros(len(items), dtype=np.uint8)
    for i, it in enumerate(items):
        flags = int(it["flags"])
        if not flags & _lib.ITEM_IS_OBJECT:
            raise ValueError(f"failed to decode response body as a {what}: element {i} is not an object")
        if mode == _lib.LIST_TABLE_ROWS and not flags & _lib.ITEM_HAS_OBJECT:
            # an empty RawExtension does not decode (responsefilterer.go:361-365)
            raise ValueError("error decoding partial object metadata from table row")
        if flags & _lib.ITEM_RAW_NAMES:  # protobuf strings: plain UTF-8, nothing to unescape
            ns = body[int(it["ns_off"]):int(it["ns_off"]) + int(it["ns_len"])].decode("utf-8", "replace")
            name = body[int(it["name_off"]):int(it["name_off"]) + int(it["name_len"])].decode("utf-8", "replace")
        else:
            ns = _text(body, int(it["ns_off"]), int(it["ns_len"]))
            name = _text(body, int(it["name_off"]), int(it["name_len"]))
        keep[i] = result.IsAllowed(ns, name)
    return _lib.list_filter(body, items, keep, ib, ie)


PROTOBUF_MEDIA_TYPE = "application/vnd.kubernetes.protobuf"


def filter_list(body: bytes, result: PrefilterResult, content_type: str = "application/json") -> bytes:
    """responsefilterer.go:376-400: keep the items whose (namespace, name) is allowed; [] when none. The body is
    decoded with the serializer its Content-Type names (responsefilterer.go:241-266): JSON, or kube's protobuf
    envelope for built-in types -- there the dropped `items` entries disappear and every other byte stays."""
    if content_type.split(";")[0].strip() == PROTOBUF_MEDIA_TYPE:
        return _filter_by_result(body, result, _lib.LIST_PROTOBUF, "list")
    return _filter_by_result(body, result, _lib.LIST_ITEMS, "list")


def filter_table(body: bytes, result: PrefilterResult) -> bytes:
    """responsefilterer.go:349-374: the same over metav1.Table rows (rows[i].object.metadata)."""
    return _filter_by_result(body, result, _lib.LIST_TABLE_ROWS, "table")


def filter_object(body: bytes, result: PrefilterResult, content_type: str = "application/json") -> bytes:
    """responsefilterer.go:320-341, :403-415: a single object passes unchanged or the request is unauthorized."""
    if content_type.split(";")[0].strip() == PROTOBUF_MEDIA_TYPE:
        try:
            scanned = _lib.list_scan(body, _lib.LIST_PROTOBUF_OBJECT)
        except _lib.ZgpuError:
            raise ValueError("failed to decode response body") from None
        ns = name = ""
        if scanned:  # an envelope without raw decodes to an empty object: no name, never allowed by name
            it = scanned[0][0]
            ns = body[int(it["ns_off"]):int(it["ns_off"]) + int(it["ns_len"])].decode("utf-8", "replace")
            name = body[int(it["name_off"]):int(it["name_off"]) + int(it["name_len"])].decode("utf-8", "replace")
        if not result.IsAllowed(ns, name):
            raise Unauthorized("unauthorized")
        return body
    wrapped = b'{"items":[' + body + b']}'  # reuse the item scanner for the one object
    try:
        scanned = _lib.list_scan(wrapped)
    except _lib.ZgpuError:
        raise ValueError("failed to decode response body") from None
    items = scanned[0] if scanned else ()
    if len(items) != 1 or not int(items[0]["flags"]) & _lib.ITEM_IS_OBJECT:
        raise ValueError("failed to decode response body")
    it = items[0]
    if not result.IsAllowed(_text(wrapped, int(it["ns_off"]), int(it["ns_len"])),
                            _text(wrapped, int(it["name_off"]), int(it["name_len"]))):
        raise Unauthorized("unauthorized")
    return body


# ---- the watch side: relationship updates -> re-checks -> allowed/denied changes ----------------------
# Reference: pkg/authz/watch.go:27-107 (RunWatch). One difference in mechanics, none in decisions: the
# reference issues one CheckPermission per update; here all updates of a WatchResponse go out as ONE
# CheckBulkPermissions (one GPU launch), and the results are reported in the same order.


@dataclass
class ResultChange:
    """watch.go:20-23."""
    allowed: bool
    namespaced_name: tuple  # (namespace, name)


def run_watch(watch_stream, check_client, rel, name_from_object_id: Callable[[str], str] = split_name,
              namespace_from_object_id: Callable[[str], object] = split_namespace) -> List[ResultChange]:
    """Drains `watch_stream` (client.WatchStream or any iterable of WatchResponse) and returns the
    ResultChange the reference would push on tracker.foundChanged for each update, in order. `rel` is the
    pre-filter's (res_type, "$", permission, subj_type, subj_id, subj_rel). A per-pair error or a failing
    bulk call ends the watch (the reference returns from RunWatch on a CheckPermission error): raises."""
    rt, _, perm, st, sid, srel = rel
    out: List[ResultChange] = []
    for resp in watch_stream:
        ids = [u.relationship.resource.object_id for u in resp.updates]  # the operation is not inspected
        if not ids:
            continue
        bulk = CheckBulkPermissionsRequest([CheckBulkPermissionsRequestItem(
            ObjectReference(rt, i), perm, SubjectReference(ObjectReference(st, sid), srel or "")) for i in ids])
        pairs = check_client.CheckBulkPermissions(bulk).pairs
        for i, pair in zip(ids, pairs):
            if pair.GetError() is not None:
                raise RuntimeError(f"error on CheckPermission: {pair.GetError()}")
            name = name_from_object_id(i)
            if not name:
                return out  # watch.go:92-94: the watch ends silently
            ns = namespace_from_object_id(i)
            out.append(ResultChange(pair.GetItem().permissionship == PERMISSIONSHIP_HAS_PERMISSION, (ns or "", name)))
    return out


# ---- the kube side of a filtered watch: frames held back until the tracker allows their object ----------
# Reference: pkg/authz/responsefilterer.go:487-714 (filterWatch). A deterministic restatement of its event loop:
# feed it what the two channels deliver (kube frames, tracker changes), collect what it would write.


class WatchFrameFilter:
    """State of one filtered watch response.

      on_frame(raw)    one frame of the kube watch stream ({"type": ..., "object": {...}}), raw bytes kept as is
      on_change(c)     one ResultChange from run_watch / the pre-filter (tracker.foundChanged)
    Both return the list of raw chunks to write to the client now.

    Mirrored decisions: a top-level v1 Status passes through and ends the stream (:585-592); only ADDED and
    MODIFIED frames are ever written, DELETED / BOOKMARK / ERROR frames are dropped (:637); an allowed object's
    frame is written at once, another one is buffered under its (namespace, name), a newer frame replacing an
    older one (:665-674); `allowed` flushes the buffered frame, `denied` forgets both (:676-694); a Table frame is
    keyed by its first row's object (:651-662); a frame that does not decode ends the stream (:578-582).
    """

    def __init__(self):
        self.allowed_names: set = set()
        self.buffered: Dict[tuple, bytes] = {}
        self.closed = False

    def on_frame(self, raw: bytes) -> List[bytes]:
        if self.closed:
            return []
        try:
            ev = json.loads(raw)
            if not isinstance(ev, dict):
                raise ValueError
        except ValueError:
            self.closed = True
            return []
        if ev.get("kind") == "Status" and ev.get("apiVersion") == "v1":
            self.closed = True
            return [raw]
        if ev.get("type") not in ("ADDED", "MODIFIED"):
            return []
        obj = ev.get("object")
        if not isinstance(obj, dict):
            return []  # "could not get object metadata": skipped
        if obj.get("kind") == "Table" and str(obj.get("apiVersion", "")).startswith("meta.k8s.io/"):
            rows = obj.get("rows") or []
            inner = rows[0].get("object") if rows and isinstance(rows[0], dict) else None
            if isinstance(inner, dict):
                obj = inner
        meta = obj.get("metadata") if isinstance(obj.get("metadata"), dict) else {}
        name = meta.get("name") if isinstance(meta.get("name"), str) else ""
        ns = meta.get("namespace") if isinstance(meta.get("namespace"), str) else ""
        key = (ns, name)
        if key in self.allowed_names:
            return [raw]
        self.buffered[key] = raw
        return []

    def on_change(self, change: ResultChange) -> List[bytes]:
        if self.closed:
            return []
        key = tuple(change.namespaced_name)
        if change.allowed:
            self.allowed_names.add(key)
            raw = self.buffered.pop(key, None)
            return [raw] if raw is not None else []
        self.allowed_names.discard(key)
        self.buffered.pop(key, None)
        return []
