/*
 * zgpu.h -- C ABI of libzgpu.so: a B200-native (sm_100a) batched Zanzibar
 * permission-check engine that drops in behind the reference proxy's
 * v1.PermissionsServiceClient boundary (authzed/spicedb-kubeapi-proxy,
 * pkg/proxy/options.go:81-82,371-377).
 *
 * Every entry point is what a cgo shim implementing v1.PermissionsServiceClient
 * would bind (INTEGRATION.md shows that shim). Plain pointers and sizes only; the
 * caller owns every buffer, the library copies what it keeps and never retains a
 * caller pointer past return; nothing throws or aborts across this boundary.
 *
 * Reference interface each group replaces (file:line in /root/reference):
 *   zg_check_bulk*              CheckBulkPermissions   pkg/authz/check.go:41-69,
 *                                                      pkg/authz/postfilter.go:127-178
 *                               CheckPermission        pkg/authz/watch.go:50-67
 *   zg_lookup_resources*        LookupResources stream pkg/authz/lookups.go:49-88
 *   zg_write_relationships      WriteRelationships     pkg/authz/distributedtx/activity.go:54-76
 *   zg_delete_relationships     DeleteRelationships    (v1 API; filter semantics update.go:207-271)
 *   zg_read_relationships       ReadRelationships      pkg/authz/distributedtx/activity.go:128-171
 *   zg_watch_read               WatchService.Watch     pkg/authz/watch.go:27-48
 *   zg_engine_create/load_schema embedded SpiceDB ctor pkg/spicedb/spicedb.go:18-57
 *
 * Return codes: 0 = OK, negative = error (message: zg_last_error).
 */
#ifndef ZGPU_H
#define ZGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZG_OK 0
#define ZG_EINVAL (-1)   /* bad argument / schema violation                      */
#define ZG_EEXIST (-2)   /* CREATE of an existing relationship                   */
#define ZG_ENOSCHEMA (-3)
#define ZG_ECUDA (-4)    /* CUDA error: the call fails closed                     */
#define ZG_EPRECOND (-5) /* a WriteRelationships precondition did not hold       */
#define ZG_ENOSNAPSHOT (-6) /* nothing published yet                             */
#define ZG_E2BIG (-7)    /* output buffer too small; required size reported      */
#define ZG_ENOMEM (-8)
#define ZG_EDEPTH (-9)   /* LookupResources: a candidate hit the dispatch-depth cap / work budget */

/* v1.CheckPermissionResponse.Permissionship (authzed-go v1.6.0); per-item codes */
#define ZG_NO_PERMISSION 1
#define ZG_HAS_PERMISSION 2
#define ZG_CONDITIONAL 3 /* never produced: caveats are out of scope              */
#define ZG_ITEM_ERROR 255 /* per-item error (depth > 50, unknown permission ...)  */

#define ZG_SREL_NONE 0xFFFFu     /* subject has no relation ("" or "...")          */
#define ZG_SREL_WILDCARD 0xFFFEu /* stored relationship whose subject is type:*   */
#define ZG_NO_OBJECT 0xFFFFFFFFu /* object id of a name that was never written    */

#define ZG_OP_TOUCH 0  /* v1.RelationshipUpdate_OPERATION_TOUCH  */
#define ZG_OP_CREATE 1 /* v1.RelationshipUpdate_OPERATION_CREATE */
#define ZG_OP_DELETE 2 /* v1.RelationshipUpdate_OPERATION_DELETE */

#define ZG_PRECOND_MUST_MATCH 1     /* v1.Precondition_OPERATION_MUST_MATCH     */
#define ZG_PRECOND_MUST_NOT_MATCH 2 /* v1.Precondition_OPERATION_MUST_NOT_MATCH */

#define ZG_MAX_DEPTH 50 /* pkg/spicedb/spicedb.go:33 WithDispatchMaxDepth(50) */

typedef struct zg_engine zg_engine;

/* zg_config.flags: build schema/store/snapshot on the host only (CPU unit tests of
 * the host logic). Every hot-path call on such an engine FAILS with ZG_ECUDA:
 * there is no CPU evaluation path in this library. */
#define ZG_FLAG_HOST_ONLY 1u
/* Forward expansion only: every direct probe is a binary search of the resource's row
 * (no direction-optimised probes from the subject's reverse rows). For A/B measurement. */
#define ZG_FLAG_FORWARD_ONLY 2u

typedef struct {
  int32_t device;            /* CUDA device ordinal; -1 = current device          */
  uint32_t flags;            /* ZG_FLAG_*                                         */
  uint64_t subquery_capacity; /* 0 = default; sub-queries buffered per pass       */
  uint32_t work_budget;      /* 0 = default; expansion rounds per 32-check batch
                                before unresolved checks report ZG_ITEM_ERROR
                                (the analogue of a request deadline)              */
  uint16_t shard_rank;       /* object-hash sharded store: this engine keeps only    */
  uint16_t shard_count;      /* relationships whose resource id % count == rank;
                                0 or 1 = whole store (replica)                       */
  uint32_t n_devices;        /* 0 / 1 = the one device above; k > 1 = devices device .. device + k - 1, each
                                holding a replica of the snapshot and answering its slice of every bulk call
                                (SURVEY.md 8e mode 1: no data-path collective); ZG_ALL_DEVICES = every
                                visible device from `device` on                                            */
  uint32_t reserved;
} zg_config;
#define ZG_ALL_DEVICES 0xFFFFFFFFu

/* One interned check: 16 bytes in, 1 byte out. Object ids are dense per type. */
typedef struct {
  uint32_t res;   /* resource object id (type implied by perm)                    */
  uint32_t subj;  /* subject object id                                            */
  uint16_t perm;  /* slot id of the permission OR relation being checked          */
  uint16_t stype; /* subject type id                                              */
  uint16_t srel;  /* subject relation slot id, or ZG_SREL_NONE                    */
  uint16_t flags; /* 0                                                            */
} zg_check;

/* One interned relationship (pkg/rules/rules.go:1050 grammar, interned). */
typedef struct {
  uint32_t res;
  uint32_t subj;  /* ignored when srel == ZG_SREL_WILDCARD                        */
  uint16_t rel;   /* slot id of the relation (implies resource type)              */
  uint16_t stype;
  uint16_t srel;  /* slot id, ZG_SREL_NONE or ZG_SREL_WILDCARD                    */
  uint16_t flags; /* 0                                                            */
} zg_tuple;

typedef struct {
  zg_tuple t;
  uint32_t expires_at; /* unix seconds, 0 = never (OptionalExpiresAt)             */
  uint32_t op;         /* ZG_OP_*                                                 */
} zg_update;

/* String forms: what the Go shim passes straight from the v1 protobuf messages. */
typedef struct {
  const char *res_type, *res_id, *relation; /* relation or permission name        */
  const char *subj_type, *subj_id, *subj_rel; /* subj_rel NULL/""/"..." = none     */
} zg_rel_str;

typedef struct {
  zg_rel_str rel;
  uint32_t expires_at;
  uint32_t op;
} zg_update_str;

/* v1.RelationshipFilter (pkg/authz/update.go:207-271): NULL/"" = unset field.   */
typedef struct {
  const char *res_type, *res_id, *relation;
  const char *subj_type, *subj_id, *subj_rel;
} zg_filter_str;

typedef struct {
  uint32_t op; /* ZG_PRECOND_* */
  zg_filter_str filter;
} zg_precondition_str;

typedef struct {
  uint64_t checks;          /* checks answered since creation                     */
  uint64_t launches;        /* kernels launched since creation                    */
  uint64_t passes;          /* sub-query passes beyond the first                  */
  uint64_t tuples;          /* relationships in the published snapshot            */
  uint64_t snapshot_bytes;  /* device bytes of the published snapshot             */
  uint64_t revision;        /* increments on every publish                        */
  uint64_t last_alg_bytes;  /* algorithmic bytes of the last counted call         */
  double last_kernel_ms;    /* device time of the last hot-path call              */
  uint64_t coalesced_launches; /* launches that answered more than one concurrent caller */
  uint64_t coalesced_requests; /* zg_check_bulk calls answered by those launches         */
  uint64_t stack_spills;       /* warp stacks that overflowed shared memory into HBM     */
  uint64_t memo_batches;       /* 32-check batches that switched the path memo on        */
  uint64_t split_batches;      /* batches answered in halves (sub-query buffer overflow) */
  uint64_t delta_publishes;    /* publishes that merged a journal of updates into the resident snapshot */
  uint64_t full_publishes;     /* publishes that rebuilt the snapshot from the relationship list        */
  double last_publish_ms;      /* device time of the last publish (merge or rebuild)                    */
  uint64_t streamed_calls;     /* host calls whose items were copied in behind the running kernel       */
  uint64_t lookup_batches;     /* launch sequences that answered more than one LookupResources           */
  uint64_t lookups_batched;    /* LookupResources calls answered by those                                */
  uint64_t devices;            /* GPUs this engine owns (counters above are summed over them)            */
} zg_stats;

/* ---- lifecycle --------------------------------------------------------- */
int zg_engine_create(const zg_config *cfg, zg_engine **out);
void zg_engine_destroy(zg_engine *e);
/* Message of the last failing call ON THE CALLING THREAD (thread-local, so concurrent callers never see
 * each other's text). A caller whose runtime migrates it between OS threads from one foreign call to the
 * next (Go) must fetch it on the thread of the failing call: pin the goroutine around the pair
 * (runtime.LockOSThread; go/gpuauthz/client.go `call`) or wrap call + zg_last_error_copy in one C
 * function of the cgo preamble. zg_last_error_copy writes at most cap - 1 bytes + NUL into buf and
 * returns the full length of the message. */
const char *zg_last_error(void);
size_t zg_last_error_copy(char *buf, size_t cap);

/* ---- schema (SpiceDB schema DSL subset; pkg/spicedb/bootstrap.yaml) ----- */
int zg_load_schema(zg_engine *e, const char *dsl, size_t len);
int zg_num_types(const zg_engine *e);
int zg_num_slots(const zg_engine *e);
int zg_type_id(const zg_engine *e, const char *type_name);            /* -1 unknown */
int zg_slot_id(const zg_engine *e, int type_id, const char *name);    /* -1 unknown */
int zg_slot_type(const zg_engine *e, int slot);
int zg_slot_is_permission(const zg_engine *e, int slot);
const char *zg_slot_name(const zg_engine *e, int slot);
const char *zg_type_name(const zg_engine *e, int type_id);

/* ---- object interning (string id <-> dense per-type u32) ---------------- */
uint32_t zg_intern_object(zg_engine *e, int type_id, const char *object_id);
uint32_t zg_find_object(const zg_engine *e, int type_id, const char *object_id); /* ZG_NO_OBJECT */
/* Copies the name into buf (NUL terminated); returns its length or ZG_E2BIG/ZG_EINVAL. */
int zg_object_name(const zg_engine *e, int type_id, uint32_t id, char *buf, size_t cap);

/* ---- relationship store -------------------------------------------------- */
/* Bulk TOUCH of interned relationships (expires may be NULL). Not visible to
 * checks until zg_publish. */
int zg_load_tuples(zg_engine *e, const zg_tuple *t, const uint32_t *expires, uint64_t n);
/* Transactional CREATE/TOUCH/DELETE; all-or-nothing. Not visible until publish. */
int zg_apply_updates(zg_engine *e, const zg_update *u, uint64_t n);
/* Builds the CSR snapshot in HBM and makes it the one every later check sees. */
int zg_publish(zg_engine *e);
/* Drops every relationship (interned names and ids stay) and publishes the empty store: what a caller
 * that mirrors another source of truth does before re-loading it (go/gpuauthz resync). The watch feed
 * restarts at the new revision. */
int zg_clear_relationships(zg_engine *e);
uint64_t zg_num_tuples(const zg_engine *e);
/* Clock used for expiration; 0 = wall clock (default). */
void zg_set_clock(zg_engine *e, int64_t unix_seconds);

/* v1 WriteRelationships: validate, check preconditions, apply, publish. The
 * write is visible to every check that starts after return (FullyConsistent). */
int zg_write_relationships(zg_engine *e, const zg_update_str *updates, uint64_t n,
                           const zg_precondition_str *pre, uint64_t n_pre);
/* v1 DeleteRelationships by filter; *n_deleted may be NULL. */
int zg_delete_relationships(zg_engine *e, const zg_filter_str *filter,
                            const zg_precondition_str *pre, uint64_t n_pre, uint64_t *n_deleted);
/* v1 ReadRelationships: matching live relationships as "type:id#rel@stype:sid[#srel]\n"
 * lines, sorted. Returns 0 or ZG_E2BIG (*need = bytes required incl. NUL). */
int zg_read_relationships(zg_engine *e, const zg_filter_str *filter, char *buf, size_t cap,
                          size_t *need, uint64_t *n_out);

/* ---- the hot path -------------------------------------------------------- */
/* CheckBulkPermissions: out[i] answers items[i] (same length, same order:
 * pkg/authz/check.go:54-57). HOST buffers; copies are inside the call. */
int zg_check_bulk(zg_engine *e, const zg_check *items, uint64_t n, uint8_t *out);
/* Same with DEVICE buffers on the given cudaStream_t (NULL = default stream);
 * asynchronous when the schema needs no sub-query pass. */
int zg_check_bulk_device(zg_engine *e, const zg_check *d_items, uint64_t n, uint8_t *d_out,
                         void *cuda_stream);
/* String form of CheckBulkPermissions / CheckPermission (n = 1): resolves, then takes the same
 * coalescing path as zg_check_bulk. */
int zg_check_bulk_str(zg_engine *e, const zg_rel_str *items, uint64_t n, uint8_t *out);
/* Only the resolution step of zg_check_bulk_str (strings -> interned checks; no GPU work, nothing is
 * interned: never-written names get the ZG_NO_OBJECT sentinels, unknown types / permissions give an
 * item that answers ZG_ITEM_ERROR). Lets a caller intern once and reuse the ids; they stay valid for
 * the life of the engine. */
int zg_resolve_checks(zg_engine *e, const zg_rel_str *items, uint64_t n, zg_check *out);
/* Packed form for a bulk request whose items share their literal fields (one rule template evaluated per
 * item: pkg/authz/postfilter.go:88-110): the n resource ids are the byte ranges
 * [res_off[i], res_off[i+1]) of res_ids (n + 1 offsets, no terminators); the subjects likewise, or with
 * subj_off == NULL ONE subject id for all items, NUL-terminated in subj_ids. Two buffers cross the
 * boundary instead of 6 n C strings. zg_check_bulk_packed = zg_resolve_checks_packed + zg_check_bulk. */
int zg_resolve_checks_packed(zg_engine *e, const char *res_type, const char *relation, const char *subj_type,
                             const char *subj_rel, const char *res_ids, const uint32_t *res_off,
                             const char *subj_ids, const uint32_t *subj_off, uint64_t n, zg_check *out);
int zg_check_bulk_packed(zg_engine *e, const char *res_type, const char *relation, const char *subj_type,
                         const char *subj_rel, const char *res_ids, const uint32_t *res_off,
                         const char *subj_ids, const uint32_t *subj_off, uint64_t n, uint8_t *out);

/* LookupResources: ids (ascending) of every object of res_type with HAS_PERMISSION.
 * Returns 0, or ZG_E2BIG with *n_out = required capacity. */
int zg_lookup_resources(zg_engine *e, uint16_t res_type, uint16_t perm, uint16_t stype,
                        uint32_t subj, uint16_t srel, uint32_t *out_ids, uint64_t cap,
                        uint64_t *n_out);
/* String form: resource object ids as '\n'-separated lines. */
int zg_lookup_resources_str(zg_engine *e, const char *res_type, const char *perm,
                            const char *subj_type, const char *subj_id, const char *subj_rel,
                            char *buf, size_t cap, size_t *need, uint64_t *n_out);

/* Pinned host memory for request/response buffers: zg_check_bulk copies straight
 * from/to such buffers (no staging memcpy). */
void *zg_host_alloc(size_t bytes);
void zg_host_free(void *p);

/* Test hook: copies row (relation slot, resource id, edge class k) of the last
 * BUILT snapshot (host copy) into out; with bit 31 of cls set, the REVERSE row of that
 * class for subject id `res`. Returns 0 / ZG_E2BIG (*n_out = size). */
int zg_debug_row(zg_engine *e, uint16_t rel_slot, uint32_t res, uint32_t cls, uint32_t *out,
                 uint64_t cap, uint64_t *n_out);

/* ---- object-hash sharded store (one engine per GPU, shard_count > 1) --------
 * A check starts on the owner of its resource; an edge to an object owned by another shard (or
 * into a non-pure permission) is raised as a sub-query. The host runs the ranks pass by pass
 * and exchanges the raised sub-queries (all-to-all) until none are left, then folds the values
 * back level by level (spicedb-kubeapi-proxy_b200/dist.py ShardedStoreChecker). On sharded
 * engines zg_check_bulk / zg_lookup_resources are refused: use these three calls. */
int zg_shard_pass(zg_engine *e, const zg_check *queries, uint64_t n, int level, uint64_t *n_sub);
/* The sub-queries pass `level` raised, in emission order (same layout; flags = hop depth). */
int zg_shard_subqueries(zg_engine *e, int level, zg_check *out, uint64_t n);
/* child_vals[i]: value of raised sub-query i (bit0 HAS, bit1 ERROR). out[q]: level 0: v1 code of
 * query q; deeper levels: its value bits, to be sent back to the rank that raised it. */
int zg_shard_fold(zg_engine *e, int level, const uint8_t *child_vals, uint64_t n_sub, uint8_t *out);

/* The same protocol with every buffer resident on the device (dist.DeviceShardedChecker): checks and raised
 * sub-queries are bucketised by owner in a kernel, exchanged device to device (NCCL all-to-all over NVLink, or
 * peer copies), and the values coming back are folded without a host copy. All pointers are DEVICE pointers of
 * the engine's GPU; every call returns after its work completed (the exchange runs on the caller's stream).
 *   zg_shard_route_dev   items (d_items, or the sub-queries raised by pass `level` when d_items is NULL) ->
 *                        d_routed in destination-major order + d_src[i] = source index of d_routed[i];
 *                        counts[d] (HOST array of n_dest) = items for destination d = res % n_dest
 *   zg_shard_pass_dev    zg_shard_pass with the level's queries already on the device
 *   zg_shard_fold_dev    d_child_vals in ROUTED order + the d_src of that routing -> d_out[q] per query of the
 *                        level (v1 codes when final_codes, else value bits for the rank that raised it)
 *   zg_shard_unroute_dev d_out[d_src[i]] = d_val[i]: answers restored to the caller's order */
int zg_shard_route_dev(zg_engine *e, const zg_check *d_items, uint64_t n, int level, uint32_t n_dest,
                       zg_check *d_routed, uint32_t *d_src, uint64_t *counts);
int zg_shard_pass_dev(zg_engine *e, const zg_check *d_queries, uint64_t n, int level, uint64_t *n_sub);
int zg_shard_fold_dev(zg_engine *e, int level, const uint8_t *d_child_vals, const uint32_t *d_src,
                      uint64_t n_sub, uint8_t *d_out, int final_codes);
int zg_shard_unroute_dev(zg_engine *e, const uint32_t *d_src, const uint8_t *d_val, uint64_t n, uint8_t *d_out);

/* Watch feed (v1.WatchServiceClient.Watch, pkg/authz/watch.go:27-48): the relationship changes made
 * visible by revisions > since_revision, oldest first, as '\n'-separated lines
 *   "<revision> <TOUCH|CREATE|DELETE> <type:id#rel@stype:sid[#srel]>[ <expires_at>]"
 * restricted to resources of res_type (NULL/"" = all; WatchRequest.OptionalObjectTypes). Only
 * zg_write_relationships / zg_delete_relationships feed it (interned bulk loads do not); a DELETE
 * of a relationship that did not exist is not a change. *through_revision = the engine's current
 * revision: pass it as the next since_revision. ZG_EPRECOND when since_revision is older than the
 * retained feed (2^20 changes), ZG_E2BIG with *need = bytes required incl. NUL. */
int zg_watch_read(zg_engine *e, uint64_t since_revision, const char *res_type, char *buf, size_t cap,
                  size_t *need, uint64_t *n_out, uint64_t *through_revision);

/* ---- list-response filter (host code, no GPU) -------
 * Replaces the unmarshal / re-marshal round trip of pkg/authz/postfilter.go:17-55
 * (filterListResponse) and the name extraction of postfilter.go:67-119: one structural pass
 * over the body, then a splice of the kept items' bytes. */
typedef struct zg_list_item {
  uint64_t begin, end;  /* byte range of the item's JSON value inside the body           */
  uint64_t name_off;    /* metadata.name: offset of the string CONTENTS (still escaped)  */
  uint64_t ns_off;      /* metadata.namespace, same                                      */
  uint32_t name_len;    /* 0 = absent or not a string                                    */
  uint32_t ns_len;
  uint32_t flags;       /* ZG_ITEM_*                                                     */
  uint32_t reserved;
} zg_list_item;
#define ZG_ITEM_IS_OBJECT 1u    /* the item is a JSON object (others are never checked)  */
#define ZG_ITEM_HAS_METADATA 2u /* ... with an object-valued "metadata"                  */
#define ZG_ITEM_HAS_OBJECT 4u   /* table rows: the row has an "object" key (any value)   */
#define ZG_ITEM_RAW_NAMES 8u    /* name / namespace ranges are plain bytes (protobuf), not escaped JSON */
#define ZG_LIST_ITEMS 0u        /* scan "items"; metadata at item level                  */
#define ZG_LIST_TABLE_ROWS 1u   /* scan "rows"; metadata under rows[i].object (metav1.Table,
                                   pkg/authz/responsefilterer.go:349-374)                */
#define ZG_LIST_PROTOBUF 2u     /* a protobuf-encoded <Kind>List (Content-Type application/vnd.kubernetes.protobuf,
                                   pkg/authz/responsefilterer.go:256-266,:301-313): magic "k8s\0", runtime.Unknown,
                                   raw = { ListMeta metadata = 1; repeated <Kind> items = 2 }. Items: begin/end =
                                   the whole `items` entry (tag, length, message), names = plain bytes;
                                   *items_begin = offset of raw's length varint, *items_end = end of raw. Only
                                   the pre-filter has a protobuf path in the reference (the post-filter
                                   json.Unmarshals the body, postfilter.go:19). */
#define ZG_LIST_PROTOBUF_OBJECT 3u /* zg_list_scan only: ONE protobuf-encoded object (a get: responsefilterer.go:320-341,
                                     filterObject :403-415): a single item = the raw payload, with its names; the
                                     caller passes the body through or answers "unauthorized" */
#define ZG_LIST_EMPTY_AS_NULL 1u /* zg_list_filter flag: nothing kept -> null, not []    */
/* Scans a kube List (or Table) body. Returns the number of elements of the top-level "items" ("rows") array
 * (0 if there is no such array: the reference then passes the body through), ZG_EINVAL on
 * malformed JSON, ZG_E2BIG if out != NULL and cap is too small. [*items_begin, *items_end) is
 * the byte range of the array, brackets included. */
int64_t zg_list_scan(const char *body, size_t len, uint32_t mode, zg_list_item *out, uint64_t cap,
                     uint64_t *items_begin, uint64_t *items_end);
/* Writes the body with only the items whose keep[i] != 0; every other byte is preserved.
 * A protobuf body (recognised by its magic) loses the dropped `items` entries and gets the length of `raw` rewritten;
 * pass the items_begin / items_end its scan returned. With nothing kept the JSON array becomes `[]` (the pre-filter's filterList / filterTable,
 * responsefilterer.go:356,377) or, with ZG_LIST_EMPTY_AS_NULL, `null` (the post-filter's re-marshal of
 * a nil slice, postfilter.go:138). Returns 0, or ZG_E2BIG with *out_len = bytes required. */
int zg_list_filter(const char *body, size_t len, const zg_list_item *items, uint64_t n,
                   const uint8_t *keep, uint64_t items_begin, uint64_t items_end, uint32_t flags,
                   char *out, size_t cap, size_t *out_len);

/* The standard post-filter template, `T:{{namespacedName}}#perm@S:subject` (or `{{name}}`), for a
 * whole scanned list at once: no per-item strings cross the boundary. */
typedef struct zg_list_template {
  const char *res_type, *permission;           /* literal fields of the template                */
  const char *subj_type, *subj_id, *subj_rel;  /* the (already resolved) subject                */
  const char *req_name, *req_namespace;        /* request fallbacks (pkg/rules/rules.go:321-326);
                                                  NULL = ""                                     */
  uint32_t id_kind;                            /* ZG_ID_NAME | ZG_ID_NAMESPACED_NAME            */
  uint32_t flags;                              /* ZG_TPL_CLEAR_NAMESPACE                        */
} zg_list_template;
#define ZG_ID_NAME 0u            /* resource id = {{name}}                                      */
#define ZG_ID_NAMESPACED_NAME 1u /* resource id = "namespace/name", or "name" without namespace  */
#define ZG_TPL_CLEAR_NAMESPACE 1u /* the request is on `namespaces` (rules.go:331-333)           */
/* Builds one interned check per scanned item straight from the body bytes (JSON escapes decoded).
 * checked[i] = 0 for items the reference never checks (not an object: postfilter.go:68-71) or whose
 * resource id comes out empty (the template does not resolve: :91-95); their out[i] is a placeholder
 * and the caller keeps them. No GPU work: feed out[] to zg_check_bulk. */
int zg_list_resolve(zg_engine *e, const char *body, size_t len, const zg_list_item *items, uint64_t n,
                    const zg_list_template *tpl, zg_check *out, uint8_t *checked);
/* Scan + resolve + ONE bulk check + splice, for a post-filter made of `n_tpl` such templates (an item
 * is kept when every template answers HAS_PERMISSION, postfilter.go:149-170). Returns 0 with the
 * filtered body in out (or the body itself, copied, when the reference would pass it through),
 * ZG_E2BIG with *out_len = bytes required, ZG_EINVAL on a malformed body. */
int zg_list_postfilter(zg_engine *e, const char *body, size_t len, const zg_list_template *tpl,
                       uint32_t n_tpl, char *out, size_t cap, size_t *out_len);

/* The pre-filter side of the same scan (pkg/authz/lookups.go:44-132 + responsefilterer.go:349-400):
 * keep[i] = 1 when the item's (namespace, name) is what an allowed resource id maps to under the usual
 * id <-> name rule ("namespace/name" -> (namespace, name); "name" -> (request namespace or "", name)).
 * `allowed` = object ids of res_type, ascending, as zg_lookup_resources returns them; self_name
 * (NULL = none) is one more allowed id given by name. mode = ZG_LIST_ITEMS | ZG_LIST_TABLE_ROWS. An
 * element that is not an object, or a table row without "object", is ZG_EINVAL (the reference fails to
 * decode such a body). No GPU work. */
int zg_list_keep_allowed(zg_engine *e, const char *body, size_t len, const zg_list_item *items, uint64_t n,
                         uint32_t mode, const char *res_type, const char *req_namespace,
                         const uint32_t *allowed, uint64_t n_allowed, const char *self_name, uint8_t *keep);
/* LookupResources (tpl's literal fields; id_kind / req_name are not used) + scan + keep + splice: the
 * whole pre-filtered List or Table response in one call. Nothing kept gives [] here. */
int zg_list_prefilter(zg_engine *e, const char *body, size_t len, uint32_t mode, const zg_list_template *tpl,
                      char *out, size_t cap, size_t *out_len);

/* ---- measurement --------------------------------------------------------- */
/* Build time and tuning macros of this library (goes into every bench line). */
const char *zg_build_info(void);
int zg_stats_get(zg_engine *e, zg_stats *out);
/* Runs the batch through the instrumented kernel variant and returns the
 * ALGORITHMIC bytes it needed (DESIGN.md "Algorithmic bytes"); not for timing. */
int zg_count_alg_bytes(zg_engine *e, const zg_check *items, uint64_t n, uint64_t *bytes);

#ifdef __cplusplus
}
#endif
#endif
