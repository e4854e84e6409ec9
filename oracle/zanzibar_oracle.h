/*
 * zanzibar_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the permission-check semantics that the reference
 * (authzed/spicedb-kubeapi-proxy) obtains from its embedded SpiceDB
 * (github.com/authzed/spicedb v1.47.1, go.mod:11 -- an un-vendored Go module that
 * is absent from /root/reference and cannot be built here: no Go toolchain).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library. The product (libzgpu.so) never links,
 * imports or calls it.
 *
 * PARITY STATUS: pinned against the reference's own golden cases G1..G14
 * (direct relation, 2-way union, nil, relation-as-permission,
 * LookupResources over those). Intersection, exclusion, arrows, userset subjects,
 * wildcards, expiration and the depth limit are "parity unpinned": no test in
 * /root/reference asserts them, so they follow SpiceDB's published semantics.
 *
 * Reference call sites whose observable behaviour this restates:
 *   pkg/authz/check.go:23-69        CheckBulkPermissions request/response contract
 *   pkg/authz/postfilter.go:97-172  CheckBulkPermissions (list post-filter)
 *   pkg/authz/lookups.go:49-88      LookupResources request / HAS_PERMISSION filter
 *   pkg/authz/watch.go:50-67        single CheckPermission
 *   pkg/spicedb/spicedb.go:25-56    engine config: depth 50, caches off, expiration on
 *   pkg/spicedb/bootstrap.yaml:1-40 schema DSL + relationship text grammar
 *   pkg/rules/rules.go:1050-1073    relationship string grammar
 *   pkg/authz/update.go:207-271     relationship filters (ReadRelationships/Delete)
 */
#ifndef ZANZIBAR_ORACLE_H
#define ZANZIBAR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zo_oracle zo_oracle;

/* v1.CheckPermissionResponse.Permissionship values (authzed-go v1.6.0) */
#define ZO_NO_PERMISSION 1
#define ZO_HAS_PERMISSION 2
#define ZO_ERROR 255

#define ZO_SREL_NONE 0xFFFFu     /* subject has no relation ("" or "...") */
#define ZO_SREL_WILDCARD 0xFFFEu /* tuple subject is type:* (store side only) */

#define ZO_OP_TOUCH 0
#define ZO_OP_CREATE 1
#define ZO_OP_DELETE 2

#define ZO_MAX_DEPTH 50 /* pkg/spicedb/spicedb.go:33 WithDispatchMaxDepth(50) */

/* Same 16-byte layout as zg_check in include/zgpu.h, on purpose: tests build one
 * array and hand it to both sides. */
typedef struct {
  uint32_t res;    /* per-type object id of the resource                      */
  uint32_t subj;   /* per-type object id of the subject                       */
  uint16_t perm;   /* slot id of the permission or relation (implies type)    */
  uint16_t stype;  /* type id of the subject                                  */
  uint16_t srel;   /* slot id of the subject relation, or ZO_SREL_NONE        */
  uint16_t flags;  /* reserved, 0                                             */
} zo_check_item;

zo_oracle *zo_create(const char *schema, char *err, size_t errlen);
void zo_destroy(zo_oracle *);

int zo_num_types(const zo_oracle *);
int zo_num_slots(const zo_oracle *);
int zo_type_id(const zo_oracle *, const char *type_name);               /* -1 unknown */
int zo_slot_id(const zo_oracle *, int type_id, const char *name);       /* -1 unknown */
int zo_slot_type(const zo_oracle *, int slot);
int zo_slot_is_permission(const zo_oracle *, int slot);
const char *zo_slot_name(const zo_oracle *, int slot);
const char *zo_type_name(const zo_oracle *, int type_id);

/* Object interning (string id <-> dense per-type u32). Numeric bulk loaders may
 * use ids that were never interned; such objects have no name. */
uint32_t zo_intern(zo_oracle *, int type_id, const char *object_id);
int64_t zo_find_object(const zo_oracle *, int type_id, const char *object_id); /* -1 */
const char *zo_object_name(const zo_oracle *, int type_id, uint32_t id);      /* NULL */

/* Mutations. expires_at: unix seconds, 0 = never. Returns 0, or <0 on error
 * (message via zo_last_error): -1 invalid (type/relation not allowed by schema),
 * -2 CREATE of an existing relationship. */
int zo_write(zo_oracle *, int op, int rel_slot, uint32_t res, int stype, uint32_t subj,
             int srel /* slot, ZO_SREL_NONE or ZO_SREL_WILDCARD */, int64_t expires_at);
/* n interned updates (zg_update layout) in one pass over the relationship log; a batch names every relationship at
 * most once. CREATE is treated as TOUCH. */
int zo_apply_batch(zo_oracle *, const void *zg_updates, uint64_t n);
/* "type:id#rel@stype:sid[#srel]" (pkg/rules/rules.go:1050) ; sid may be "*" */
int zo_write_str(zo_oracle *, int op, const char *rel, int64_t expires_at);
/* Bulk TOUCH of n relationships sharing (rel_slot, stype, srel). */
int zo_add_bulk(zo_oracle *, int rel_slot, int stype, int srel, const uint32_t *res,
                const uint32_t *subj, uint64_t n);
uint64_t zo_num_tuples(const zo_oracle *);

/* Check. now: unix seconds used for expiration. Returns ZO_NO_PERMISSION,
 * ZO_HAS_PERMISSION or ZO_ERROR (depth > 50 on a path that decides the answer,
 * unknown slot). */
int zo_check(zo_oracle *, const zo_check_item *item, int64_t now);
int zo_check_str(zo_oracle *, const char *res_type, const char *res_id, const char *perm,
                 const char *subj_type, const char *subj_id, const char *subj_rel,
                 int64_t now);
/* nthreads <= 0: all online cores. */
int zo_check_bulk(zo_oracle *, const zo_check_item *items, uint64_t n, uint8_t *out,
                  int nthreads, int64_t now);
/* SURVEY.md 8(d) canonical forward-evaluation byte count B(q), no short circuit. */
uint64_t zo_check_bytes(zo_oracle *, const zo_check_item *items, uint64_t n, int64_t now);

/* LookupResources by definition: every object of res_type that is the resource
 * of at least one live relationship and for which Check == HAS. Sorted ids.
 * Returns 0, or -7 (E2BIG) with *n_out = required size. */
int zo_lookup_resources(zo_oracle *, int res_type, int perm_slot, int stype, uint32_t subj,
                        int srel, int64_t now, uint32_t *out_ids, uint64_t cap,
                        uint64_t *n_out);

/* ReadRelationships-style filter (pkg/authz/update.go:207-271). Any field < 0 /
 * NULL = unset. Emits matching live relationships as text lines into buf
 * ("type:id#rel@stype:sid[#srel]\n", sorted); returns count or -7 if cap too small
 * (*need = required bytes). */
int64_t zo_read_str(zo_oracle *, const char *res_type, const char *res_id, const char *rel,
                    const char *subj_type, const char *subj_id, const char *subj_rel,
                    int64_t now, char *buf, size_t cap, size_t *need);

const char *zo_last_error(const zo_oracle *);

#ifdef __cplusplus
}
#endif
#endif
