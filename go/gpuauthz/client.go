// Package gpuauthz is the reference-side binding of libzgpu: an implementation of
// v1.PermissionsServiceClient (authzed-go v1.6.0) that answers CheckPermission,
// CheckBulkPermissions and LookupResources on a B200 through the C ABI in
// include/zgpu.h, and delegates everything else to the embedded SpiceDB it wraps.
//
// Injection point in the reference: pkg/proxy/options.go:371-377 only assigns
// Options.PermissionsClient when it is nil, so
//
//	opts.PermissionsClient = gpuauthz.New(embeddedClient, schemaText)
//
// before opts.Complete() is all the proxy needs (pkg/proxy/server.go:136-139,153 pass
// the same client to the workflow activities and to pkg/authz).
//
// STATUS: written against the v1 API from the call sites in the reference
// (pkg/authz/check.go:23-69, lookups.go:49-88, watch.go:50-67, postfilter.go:97-172,
// distributedtx/activity.go:54-171). It has NOT been compiled: there is no Go toolchain
// in the build image or on the GPU box. The Python mirror
// (spicedb-kubeapi-proxy_b200/client.py) drives the identical C ABI in the tests.
package gpuauthz

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../spicedb-kubeapi-proxy_b200 -lzgpu -Wl,-rpath,${SRCDIR}/../../spicedb-kubeapi-proxy_b200
#include <stdlib.h>
#include <string.h>
#include "zgpu.h"

static zg_rel_str *rel_array(size_t n) { return (zg_rel_str *)calloc(n ? n : 1, sizeof(zg_rel_str)); }
static zg_update_str *upd_array(size_t n) { return (zg_update_str *)calloc(n ? n : 1, sizeof(zg_update_str)); }
static zg_precondition_str *pre_array(size_t n) { return (zg_precondition_str *)calloc(n ? n : 1, sizeof(zg_precondition_str)); }
*/
import "C"

import (
	"context"
	"fmt"
	"io"
	"runtime"
	"strings"
	"sync"
	"sync/atomic"
	"unsafe"

	v1 "github.com/authzed/authzed-go/proto/authzed/api/v1"
	"google.golang.org/grpc"
	"google.golang.org/grpc/codes"
	"google.golang.org/grpc/metadata"
	"google.golang.org/grpc/status"
)

// Client implements v1.PermissionsServiceClient.
type Client struct {
	v1.PermissionsServiceClient // the embedded SpiceDB: source of truth for everything not overridden

	engine *C.zg_engine
	mu     sync.Mutex // serialises the write mirror; checks are thread-safe in the library
	// broken is set when the mirror diverged from SpiceDB and could not be rebuilt: every check path
	// then fails (pkg/authz/authz.go:93-97 turns an error into a denial) instead of answering from a
	// store that may still hold revoked grants.
	broken atomic.Bool
}

// AllDevices: every visible GPU from `device` on (zg_config.n_devices).
const AllDevices = uint32(C.ZG_ALL_DEVICES)

var errBroken = status.Error(codes.Unavailable, "gpu mirror is out of sync with SpiceDB and could not be rebuilt")

// New creates the GPU engine, loads the schema and mirrors the relationships that are
// already in SpiceDB (bootstrap file: pkg/spicedb/spicedb.go:19-24).
//
// nDevices > 1 (or AllDevices) makes the one engine own that many GPUs of the box, each holding a replica of the
// snapshot and answering its slice of every bulk call: the proxy keeps a single client (pkg/proxy/options.go:81-82),
// so this is how it uses more than one GPU.
func New(ctx context.Context, inner v1.PermissionsServiceClient, schema string, device int, nDevices uint32) (*Client, error) {
	cfg := C.zg_config{device: C.int32_t(device), n_devices: C.uint32_t(nDevices)}
	var eng *C.zg_engine
	if err := call(func() C.int { return C.zg_engine_create(&cfg, &eng) }); err != nil {
		return nil, err // fail closed: no CPU fallback inside the library
	}
	c := &Client{PermissionsServiceClient: inner, engine: eng}
	cs := C.CString(schema)
	defer C.free(unsafe.Pointer(cs))
	if err := call(func() C.int { return C.zg_load_schema(eng, cs, C.size_t(len(schema))) }); err != nil {
		c.Close()
		return nil, err
	}
	if err := c.resync(ctx); err != nil {
		c.Close()
		return nil, err
	}
	return c, nil
}

func (c *Client) Close() {
	if c.engine != nil {
		C.zg_engine_destroy(c.engine)
		c.engine = nil
	}
}

// call runs one C entry point and, when it fails, fetches the library's message. zg_last_error is
// thread-local; a goroutine may migrate between OS threads from one cgo call to the next, so the pair
// (failing call, zg_last_error) runs with the goroutine pinned to its thread.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != 0 {
		return lastError(rc)
	}
	return nil
}

// lastError must run on the OS thread of the failing call: only use it through call().
func lastError(rc C.int) error {
	code := codes.Internal
	switch rc {
	case C.ZG_EINVAL:
		code = codes.InvalidArgument
	case C.ZG_EEXIST:
		code = codes.AlreadyExists
	case C.ZG_EPRECOND, C.ZG_ENOSCHEMA, C.ZG_ENOSNAPSHOT:
		code = codes.FailedPrecondition
	case C.ZG_ENOMEM, C.ZG_EDEPTH:
		code = codes.ResourceExhausted
	}
	return status.Error(code, C.GoString(C.zg_last_error()))
}

// cstrs owns the C strings of one call.
type cstrs struct{ p []unsafe.Pointer }

func (s *cstrs) add(v string) *C.char {
	c := C.CString(v)
	s.p = append(s.p, unsafe.Pointer(c))
	return c
}
func (s *cstrs) free() {
	for _, p := range s.p {
		C.free(p)
	}
}

func (s *cstrs) fillRel(dst *C.zg_rel_str, res *v1.ObjectReference, relation string, subj *v1.SubjectReference) {
	dst.res_type = s.add(res.GetObjectType())
	dst.res_id = s.add(res.GetObjectId())
	dst.relation = s.add(relation)
	dst.subj_type = s.add(subj.GetObject().GetObjectType())
	dst.subj_id = s.add(subj.GetObject().GetObjectId())
	dst.subj_rel = s.add(subj.GetOptionalRelation())
}

func (s *cstrs) fillFilter(dst *C.zg_filter_str, f *v1.RelationshipFilter) {
	dst.res_type = s.add(f.GetResourceType())
	dst.res_id = s.add(f.GetOptionalResourceId())
	dst.relation = s.add(f.GetOptionalRelation())
	if sf := f.GetOptionalSubjectFilter(); sf != nil {
		dst.subj_type = s.add(sf.GetSubjectType())
		dst.subj_id = s.add(sf.GetOptionalSubjectId())
		if r := sf.GetOptionalRelation(); r != nil {
			rel := r.GetRelation()
			if rel == "" {
				rel = "..."
			}
			dst.subj_rel = s.add(rel)
		}
	}
}

// CheckBulkPermissions: pkg/authz/check.go:41-69 and postfilter.go:127-178 rely on
// Pairs[i] answering Items[i], and only on == PERMISSIONSHIP_HAS_PERMISSION.
func (c *Client) CheckBulkPermissions(ctx context.Context, req *v1.CheckBulkPermissionsRequest, _ ...grpc.CallOption) (*v1.CheckBulkPermissionsResponse, error) {
	n := len(req.GetItems())
	resp := &v1.CheckBulkPermissionsResponse{Pairs: make([]*v1.CheckBulkPermissionsPair, n)}
	if n == 0 {
		return resp, nil
	}
	var s cstrs
	defer s.free()
	arr := C.rel_array(C.size_t(n))
	defer C.free(unsafe.Pointer(arr))
	items := unsafe.Slice(arr, n)
	for i, it := range req.GetItems() {
		s.fillRel(&items[i], it.GetResource(), it.GetPermission(), it.GetSubject())
	}
	out := (*C.uint8_t)(C.malloc(C.size_t(n)))
	defer C.free(unsafe.Pointer(out))
	if c.broken.Load() {
		return nil, errBroken
	}
	if err := call(func() C.int { return C.zg_check_bulk_str(c.engine, arr, C.uint64_t(n), out) }); err != nil {
		return nil, err // fail closed: pkg/authz/authz.go:93-97 turns any error into a denial
	}
	codesOut := unsafe.Slice((*byte)(unsafe.Pointer(out)), n)
	for i, it := range req.GetItems() {
		pair := &v1.CheckBulkPermissionsPair{Request: it}
		if codesOut[i] == C.ZG_ITEM_ERROR {
			pair.Response = &v1.CheckBulkPermissionsPair_Error{Error: status.New(codes.FailedPrecondition,
				"check failed: unknown permission or type, or maximum depth exceeded").Proto()}
		} else {
			pair.Response = &v1.CheckBulkPermissionsPair_Item{Item: &v1.CheckBulkPermissionsResponseItem{
				Permissionship: v1.CheckPermissionResponse_Permissionship(codesOut[i])}}
		}
		resp.Pairs[i] = pair
	}
	return resp, nil
}

// CheckPermission: pkg/authz/watch.go:50-67 (one call per relationship update).
func (c *Client) CheckPermission(ctx context.Context, req *v1.CheckPermissionRequest, opts ...grpc.CallOption) (*v1.CheckPermissionResponse, error) {
	bulk, err := c.CheckBulkPermissions(ctx, &v1.CheckBulkPermissionsRequest{
		Consistency: req.GetConsistency(),
		Items: []*v1.CheckBulkPermissionsRequestItem{{
			Resource: req.GetResource(), Permission: req.GetPermission(), Subject: req.GetSubject()}},
	}, opts...)
	if err != nil {
		return nil, err
	}
	if e := bulk.Pairs[0].GetError(); e != nil {
		return nil, status.ErrorProto(e)
	}
	return &v1.CheckPermissionResponse{Permissionship: bulk.Pairs[0].GetItem().GetPermissionship()}, nil
}

// LookupResources: pkg/authz/lookups.go:65-88 reads the stream until io.EOF and keeps
// LOOKUP_PERMISSIONSHIP_HAS_PERMISSION entries; order is irrelevant (a set).
func (c *Client) LookupResources(ctx context.Context, req *v1.LookupResourcesRequest, _ ...grpc.CallOption) (v1.PermissionsService_LookupResourcesClient, error) {
	if c.broken.Load() {
		return nil, errBroken
	}
	var s cstrs
	defer s.free()
	subj := req.GetSubject()
	capacity := C.size_t(1 << 16)
	for {
		buf := (*C.char)(C.malloc(capacity))
		var need C.size_t
		var n C.uint64_t
		var rc C.int
		err := call(func() C.int {
			rc = C.zg_lookup_resources_str(c.engine, s.add(req.GetResourceObjectType()), s.add(req.GetPermission()),
				s.add(subj.GetObject().GetObjectType()), s.add(subj.GetObject().GetObjectId()),
				s.add(subj.GetOptionalRelation()), buf, capacity, &need, &n)
			if rc == C.ZG_E2BIG {
				return 0
			}
			return rc
		})
		if rc == C.ZG_E2BIG {
			C.free(unsafe.Pointer(buf))
			capacity = need + 16
			continue
		}
		if err != nil {
			C.free(unsafe.Pointer(buf))
			return nil, err
		}
		ids := strings.Split(strings.TrimSuffix(C.GoString(buf), "\n"), "\n")
		C.free(unsafe.Pointer(buf))
		if n == 0 {
			ids = nil
		}
		return &sliceStream{ctx: ctx, ids: ids}, nil
	}
}

type sliceStream struct {
	ctx context.Context
	ids []string
	pos int
}

func (s *sliceStream) Recv() (*v1.LookupResourcesResponse, error) {
	if err := s.ctx.Err(); err != nil {
		return nil, status.FromContextError(err).Err() // codes.Canceled is special-cased at responsefilterer.go:170
	}
	if s.pos >= len(s.ids) {
		return nil, io.EOF
	}
	id := s.ids[s.pos]
	s.pos++
	return &v1.LookupResourcesResponse{ResourceObjectId: id,
		Permissionship: v1.LookupPermissionship_LOOKUP_PERMISSIONSHIP_HAS_PERMISSION}, nil
}
func (s *sliceStream) Header() (metadata.MD, error) { return nil, nil }
func (s *sliceStream) Trailer() metadata.MD         { return nil }
func (s *sliceStream) CloseSend() error             { return nil }
func (s *sliceStream) Context() context.Context     { return s.ctx }
func (s *sliceStream) SendMsg(any) error            { return nil }
func (s *sliceStream) RecvMsg(any) error            { return io.EOF }

// WriteRelationships: SpiceDB stays the source of truth (preconditions, idempotency keys
// with expiry: pkg/authz/distributedtx/activity.go:54-102); a successful write is mirrored
// into the GPU store before returning so FullyConsistent checks observe it.
func (c *Client) WriteRelationships(ctx context.Context, req *v1.WriteRelationshipsRequest, opts ...grpc.CallOption) (*v1.WriteRelationshipsResponse, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	resp, err := c.PermissionsServiceClient.WriteRelationships(ctx, req, opts...)
	if err != nil {
		return nil, err
	}
	if err := c.mirror(req.GetUpdates()); err != nil {
		// the stores diverged (SpiceDB committed, the mirror rejected the batch -- DELETEs included):
		// rebuild the mirror from SpiceDB rather than serve stale answers
		if rerr := c.resync(ctx); rerr != nil {
			c.broken.Store(true)
			return nil, status.Errorf(codes.Internal, "gpu mirror failed (%v) and resync failed: %v", err, rerr)
		}
	}
	return resp, nil
}

func (c *Client) mirror(updates []*v1.RelationshipUpdate) error {
	n := len(updates)
	if n == 0 {
		return nil
	}
	var s cstrs
	defer s.free()
	arr := C.upd_array(C.size_t(n))
	defer C.free(unsafe.Pointer(arr))
	ups := unsafe.Slice(arr, n)
	for i, u := range updates {
		r := u.GetRelationship()
		s.fillRel(&ups[i].rel, r.GetResource(), r.GetRelation(), r.GetSubject())
		if ts := r.GetOptionalExpiresAt(); ts != nil {
			ups[i].expires_at = C.uint32_t(ts.GetSeconds())
		}
		switch u.GetOperation() {
		case v1.RelationshipUpdate_OPERATION_CREATE, v1.RelationshipUpdate_OPERATION_TOUCH:
			ups[i].op = C.ZG_OP_TOUCH // SpiceDB already enforced CREATE semantics
		case v1.RelationshipUpdate_OPERATION_DELETE:
			ups[i].op = C.ZG_OP_DELETE
		}
	}
	return call(func() C.int { return C.zg_write_relationships(c.engine, arr, C.uint64_t(n), nil, 0) })
}

// DeleteRelationships: delegate, then apply the same filter to the mirror.
func (c *Client) DeleteRelationships(ctx context.Context, req *v1.DeleteRelationshipsRequest, opts ...grpc.CallOption) (*v1.DeleteRelationshipsResponse, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	resp, err := c.PermissionsServiceClient.DeleteRelationships(ctx, req, opts...)
	if err != nil {
		return nil, err
	}
	var s cstrs
	defer s.free()
	var f C.zg_filter_str
	s.fillFilter(&f, req.GetRelationshipFilter())
	if err := call(func() C.int { return C.zg_delete_relationships(c.engine, &f, nil, 0, nil) }); err != nil {
		if rerr := c.resync(ctx); rerr != nil {
			c.broken.Store(true)
			return nil, status.Errorf(codes.Internal, "gpu mirror delete failed (%v) and resync failed: %v", err, rerr)
		}
	}
	return resp, nil
}

// resync rebuilds the mirror from a full ReadRelationships per resource type. The mirror is EMPTIED
// first: it runs when a mirrored write failed after SpiceDB committed, and a batch that was rejected as a
// whole may have carried DELETEs -- re-TOUCHing what SpiceDB holds now would leave those revoked grants
// in place (fail-open). Until the rebuilt snapshot is published nothing is visible: zg_clear_relationships
// publishes the empty store, so a check that races the rebuild is denied, not wrongly allowed.
func (c *Client) resync(ctx context.Context) error {
	if err := call(func() C.int { return C.zg_clear_relationships(c.engine) }); err != nil {
		return err
	}
	nTypes := int(C.zg_num_types(c.engine))
	for t := 0; t < nTypes; t++ {
		typeName := C.GoString(C.zg_type_name(c.engine, C.int(t)))
		stream, err := c.PermissionsServiceClient.ReadRelationships(ctx, &v1.ReadRelationshipsRequest{
			Consistency:        &v1.Consistency{Requirement: &v1.Consistency_FullyConsistent{FullyConsistent: true}},
			RelationshipFilter: &v1.RelationshipFilter{ResourceType: typeName},
		})
		if err != nil {
			return err
		}
		var batch []*v1.RelationshipUpdate
		flush := func() error {
			if len(batch) == 0 {
				return nil
			}
			err := c.mirror(batch)
			batch = batch[:0]
			return err
		}
		for {
			r, err := stream.Recv()
			if err == io.EOF {
				break
			}
			if err != nil {
				return err
			}
			batch = append(batch, &v1.RelationshipUpdate{Operation: v1.RelationshipUpdate_OPERATION_TOUCH, Relationship: r.GetRelationship()})
			if len(batch) == 1000 { // pkg/spicedb/spicedb.go:34 WithMaximumUpdatesPerWrite(1000)
				if err := flush(); err != nil {
					return err
				}
			}
		}
		if err := flush(); err != nil {
			return err
		}
	}
	if err := call(func() C.int { return C.zg_publish(c.engine) }); err != nil {
		return err
	}
	c.broken.Store(false)
	return nil
}

// ListTemplate is the standard post-filter template, "T:{{namespacedName}}#perm@S:subject"
// (or {{name}}), with its literal fields already resolved for this request.
type ListTemplate struct {
	ResourceType, Permission                string
	SubjectType, SubjectID, SubjectRelation string
	RequestName, RequestNamespace           string // fallbacks of rules.NewResolveInput (pkg/rules/rules.go:321-326)
	NameOnly                                bool   // {{name}} instead of {{namespacedName}}
	ClearNamespace                          bool   // the request is on `namespaces` (rules.go:331-333)
}

// FilterListResponse replaces the body of filterListResponse (pkg/authz/postfilter.go:17-55) for rules whose
// PostFilters are all standard templates: one structural scan of the body, one bulk check on the GPU, one
// splice of the kept items -- no map[string]interface{} round trip, no per-item strings across cgo.
// The result is the same JSON value the reference produces (unknown fields, key order and number spelling
// are the apiserver's, since nothing is re-serialised).
func (c *Client) FilterListResponse(body []byte, tpls []ListTemplate) ([]byte, error) {
	if len(body) == 0 {
		return nil, fmt.Errorf("failed to parse list response: empty body")
	}
	var cs cstrs
	defer cs.free()
	ct := make([]C.zg_list_template, len(tpls))
	for i, t := range tpls {
		ct[i] = C.zg_list_template{
			res_type: cs.add(t.ResourceType), permission: cs.add(t.Permission),
			subj_type: cs.add(t.SubjectType), subj_id: cs.add(t.SubjectID), subj_rel: cs.add(t.SubjectRelation),
			req_name: cs.add(t.RequestName), req_namespace: cs.add(t.RequestNamespace),
			id_kind: C.ZG_ID_NAMESPACED_NAME,
		}
		if t.NameOnly {
			ct[i].id_kind = C.ZG_ID_NAME
		}
		if t.ClearNamespace {
			ct[i].flags = C.ZG_TPL_CLEAR_NAMESPACE
		}
	}
	var tp *C.zg_list_template
	if len(ct) > 0 {
		tp = &ct[0]
	}
	out := make([]byte, len(body)+8) // the filtered body is never longer than body + 2
	var outLen C.size_t
	if c.broken.Load() {
		return nil, errBroken
	}
	if err := call(func() C.int {
		return C.zg_list_postfilter(c.engine, (*C.char)(unsafe.Pointer(&body[0])), C.size_t(len(body)), tp, C.uint32_t(len(ct)),
			(*C.char)(unsafe.Pointer(&out[0])), C.size_t(len(out)), &outLen)
	}); err != nil {
		return nil, fmt.Errorf("failed to filter items with bulk permissions: %w", err)
	}
	return out[:outLen], nil
}

// PrefilterListResponse replaces runLookupResources + filterList / filterTable (pkg/authz/lookups.go:44-132,
// pkg/authz/responsefilterer.go:349-400) for rules whose pre-filter is the usual "namespace/name" split: one
// LookupResources on the GPU (concurrent list requests share launches), one scan of the body, one splice.
// contentType is the response's Content-Type header: kube clients get protobuf for built-in types and the reference
// decodes with the serializer it names (responsefilterer.go:241-266); asTable = the request's Accept carried
// "as=Table" (always JSON, responsefilterer.go:344-346).
func (c *Client) PrefilterListResponse(body []byte, contentType string, asTable bool, t ListTemplate) ([]byte, error) {
	if len(body) == 0 {
		return nil, fmt.Errorf("failed to decode response body: empty body")
	}
	mode := C.uint32_t(C.ZG_LIST_ITEMS)
	switch {
	case asTable:
		mode = C.ZG_LIST_TABLE_ROWS
	case strings.HasPrefix(strings.TrimSpace(contentType), "application/vnd.kubernetes.protobuf"):
		mode = C.ZG_LIST_PROTOBUF
	}
	var cs cstrs
	defer cs.free()
	ct := C.zg_list_template{
		res_type: cs.add(t.ResourceType), permission: cs.add(t.Permission),
		subj_type: cs.add(t.SubjectType), subj_id: cs.add(t.SubjectID), subj_rel: cs.add(t.SubjectRelation),
		req_name: cs.add(t.RequestName), req_namespace: cs.add(t.RequestNamespace),
	}
	out := make([]byte, len(body)+8)
	var outLen C.size_t
	if c.broken.Load() {
		return nil, errBroken
	}
	if err := call(func() C.int {
		return C.zg_list_prefilter(c.engine, (*C.char)(unsafe.Pointer(&body[0])), C.size_t(len(body)), mode, &ct,
			(*C.char)(unsafe.Pointer(&out[0])), C.size_t(len(out)), &outLen)
	}); err != nil {
		return nil, fmt.Errorf("failed to filter response: %w", err)
	}
	return out[:outLen], nil
}

var _ v1.PermissionsServiceClient = (*Client)(nil)
var _ = fmt.Sprintf
