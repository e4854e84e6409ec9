// schema.h -- SpiceDB schema DSL subset -> flat device program.
//
// The reference configures its engine with schema text (pkg/spicedb/bootstrap.yaml:1-38,
// pkg/spicedb/spicedb.go:19-24). This compiler turns that text into the tables the
// CUDA kernels walk:
//   * every relation becomes a "data relation" with one EDGE CLASS per allowed
//     subject kind (type, type#rel, type:*): subject type and child slot are static
//     per class, so an edge is a bare u32 object id;
//   * every permission whose (same-object-inlined) expression is a pure union
//     becomes a UNIT: a short list of REL / ARROW ops evaluated as reachability;
//   * permissions containing & or - become a TREE: a postfix boolean program over
//     leaf units, evaluated after the reachability pass (DESIGN.md "Boolean
//     structure").
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace zg {

constexpr uint16_t kNone = 0xFFFF;      // ZG_SREL_NONE
constexpr uint16_t kWildcard = 0xFFFE;  // ZG_SREL_WILDCARD
constexpr int kMaxClasses = 8;          // edge classes per relation
constexpr int kMaxLeaves = 32;          // leaf units per non-pure permission

// ---- device-visible PODs (copied verbatim into the program blob) -------------
enum SlotKind : uint16_t { SK_RELATION = 0, SK_PURE = 1, SK_NONPURE = 2 };
enum OpKind : uint16_t { OP_REL = 0, OP_ARROW = 1 };
enum TreeOpKind : uint16_t { T_LEAF = 0, T_TRIVIAL = 1, T_OR = 2, T_AND = 3, T_ANDNOT = 4 };
enum ClsFlags : uint16_t {
  CF_EXPIRY = 1,  // relationships of this class may carry an expiration
  CF_EMPTY = 2,   // no relationship of this class in the whole snapshot (set at publish)
  CF_INVERT = 4,  // direct class whose probes may be answered from the subject's reverse row
};
enum UnitFlags : uint16_t {
  UF_EXPANSIVE = 1,  // unit can push userset / arrow ranges
  UF_LEAF_INV = 2,   // in this snapshot every step is an invertible direct probe: a visit can
                     // only answer "is (class, obj) one of the subject's own memberships"
};
constexpr uint16_t kStepTargetLeaf = 0x100;  // DStep.flags: children are UF_LEAF_INV visits
constexpr uint16_t kStepTargetL2 = 0x200;    // DStep.flags: the children's unit is ONE userset edge class whose own
                                             // children are UF_LEAF_INV visits (resource -> team#member -> user): the
                                             // whole two-level range is answered from the subject's reverse rows

struct DSlot {   // per slot (relation or permission)
  uint16_t kind;        // SlotKind
  uint16_t unit;        // SK_RELATION / SK_PURE: unit id ; SK_NONPURE: tree id
  uint16_t type;
  uint16_t pad;
};
struct DUnit {   // pure-union program evaluated at one object
  uint16_t op_begin, op_end;    // into ops[]
  uint16_t mem_begin, mem_end;  // into members[]: slots inlined into this unit
  uint16_t flags;               // UnitFlags
  uint16_t step_begin, step_end;  // into steps[]: the unit flattened for this snapshot, probes first:
  uint16_t push_begin;            // [step_begin, push_begin) DIRECT / WILD (lane-local), [push_begin, step_end) PUSH
};
// One edge class a unit touches, with everything the visit needs in one 24-byte record.
// Built at publish: classes that are empty in the snapshot produce no step at all.
enum StepKind : uint16_t { ST_DIRECT = 0, ST_WILD = 1, ST_PUSH = 2 };
struct DStep {
  uint64_t row_base;  // row_ptr index of (object 0, this class): rel.row_base + k
  uint32_t nres;      // objects covered by the relation's row table
  uint16_t ncls;      // row stride per object (all classes of the resource type)
  uint16_t kind;      // StepKind
  uint16_t stype;     // ST_DIRECT / ST_WILD: subject type that can match
  uint16_t tslot;     // ST_PUSH: slot the children are visited at
  uint16_t flags;     // ClsFlags (CF_EXPIRY, CF_INVERT) | kStepTargetLeaf
  uint16_t gc;        // global class id (key of the subject's reverse-row set)
  uint16_t tunit;     // ST_PUSH: unit of tslot (kNone when tslot is a non-pure permission)
  uint16_t tgc;       // kStepTargetLeaf: the single class of the target unit; kStepTargetL2: the single
                      // (userset) class of the target unit, whose reverse rows are streamed
  uint16_t tinv;      // ST_DIRECT: index of gc among its subject type's invertible classes;
                      // kStepTargetLeaf: the same for tgc; kStepTargetL2: for the innermost direct class
  uint16_t tstype;    // kStepTargetLeaf / kStepTargetL2: subject type the (innermost) direct class accepts
};
struct DOp {
  uint16_t kind;       // OpKind
  uint16_t rel;        // data relation index
  uint16_t tgt_begin;  // OP_ARROW: tgts[tgt_begin + class] = child slot or kNone
  uint16_t pad;
};
struct DRel {    // data relation: rows of (object x class). All relations of one resource type
                 // are interleaved per object, so one object's offsets share a sector or two.
  uint64_t row_base;   // index into row_ptr pool of (object 0, class 0 of this relation)
  uint32_t nres;       // objects covered
  uint16_t ncls;       // classes of this relation
  uint16_t cls_begin;  // into classes[]
  uint32_t stride;     // classes of ALL relations of the resource type = row stride per object
  uint32_t pad;
};
struct DCls {
  uint16_t stype;
  uint16_t sslot;      // kNone (direct), kWildcard, or subject relation slot
  uint16_t flags;      // ClsFlags
  uint16_t rel;        // data relation index this class belongs to
  uint32_t nsubj;      // reverse CSR: subject objects covered (1 for a wildcard class)
  uint16_t rtype;      // type of the resources of this class
  uint16_t rstride;    // reverse CSR row stride per subject: the reverse rows of ALL classes of one
                       // subject type are interleaved per subject (one or two sectors hold every
                       // offset of a subject), like the forward tables are per object
  uint64_t rrow_base;  // reverse CSR: rrow_ptr index of (subject 0, this class); row of subject s =
                       // rrow_base + s * rstride
};
struct DTypeInv {  // per subject type: the invertible direct classes (inv_cls[begin, end))
  uint16_t begin, end;
};
struct DTree {   // postfix boolean program of a non-pure permission
  uint16_t op_begin, op_end;     // into tree_ops[]
  uint16_t leaf_begin, n_leaves; // leaf_units[leaf_begin + i] = unit id
  uint16_t flags;                // UF_EXPANSIVE when any leaf unit is
  uint16_t pad;
};
struct DTreeOp {
  uint16_t kind;  // TreeOpKind
  uint16_t arg;   // T_LEAF: leaf index ; T_TRIVIAL: slot
};
struct DHeader {  // first bytes of the blob; offsets in bytes from blob start
  uint32_t magic, total_bytes;
  uint32_t n_types, n_slots, n_units, n_ops, n_rels, n_cls, n_trees, n_tree_ops;
  uint32_t off_slots, off_units, off_ops, off_rels, off_cls, off_tgts, off_members;
  uint32_t off_trees, off_tree_ops, off_leaf_units, off_reach;
  uint32_t off_type_inv, off_inv_cls, off_steps, n_steps;
  uint32_t off_type_rcls, off_rcls;  // per subject type: every class whose subjects have that type
  uint32_t max_leaves;   // job stride L in general mode (1 if no non-pure slot)
  uint32_t has_nonpure;
  uint32_t has_expiry;
  uint32_t reach_words;  // u32 words per slot in reach[] (bit t: subjects of type t reachable)
};

// ---- host-side model -----------------------------------------------------------
struct ClassInfo {
  uint16_t stype;
  uint16_t sslot;
  bool expiry;
};
struct Expr {
  enum Kind { NIL, REF, ARROW, UNION, INTER, EXCL } kind = NIL;
  int slot = -1;   // REF: slot ; ARROW: tupleset relation slot
  int name = -1;   // ARROW: name id of the computed permission
  int l = -1, r = -1;
};
struct SlotInfo {
  std::string name;
  int name_id = -1;
  uint16_t type = 0;
  bool is_perm = false;
  std::vector<ClassInfo> classes;  // relations
  int expr = -1;                   // permissions: root in Schema::exprs
  int rel_index = -1;              // relations: data relation index
  uint16_t kind = SK_RELATION;
  int unit = -1, tree = -1;
};
struct TypeInfo {
  std::string name;
  std::vector<int> slots;
};

class Schema {
 public:
  // Returns empty string on success, else the error message.
  std::string parse(const std::string& text);

  int type_id(const std::string& n) const;
  int slot_id(int type, const std::string& n) const;
  int slot_by_name_id(int type, int name_id) const;
  int class_of(int rel_slot, uint16_t stype, uint16_t sslot) const;  // -1 if not allowed

  std::vector<TypeInfo> types;
  std::vector<SlotInfo> slots;
  std::vector<std::string> names;
  std::vector<Expr> exprs;
  std::vector<int> rel_slots;  // data relation index -> slot
  bool use_expiration = false;

  // compiled tables
  std::vector<DSlot> d_slots;
  std::vector<DUnit> d_units;
  std::vector<DOp> d_ops;
  std::vector<DCls> d_cls;
  std::vector<uint16_t> d_tgts, d_members, d_leaf_units;
  std::vector<DTree> d_trees;
  std::vector<DTreeOp> d_tree_ops;
  std::vector<uint32_t> d_reach;
  uint32_t reach_words = 1;
  uint32_t max_leaves = 1;
  bool has_nonpure = false, has_expiry = false;

  // Serialises the program; DRel rows (row_base, nres) come from the store.
  // cls: per-class dynamic data from the store (rrow_base, CF_EMPTY), same order as d_cls.
  std::vector<uint8_t> blob(const std::vector<DRel>& rels, const std::vector<DCls>& cls) const;
  std::vector<DTypeInv> d_type_inv, d_type_rcls;
  std::vector<uint16_t> d_inv_cls, d_rcls;

 private:
  std::string compile();
  int name_id(const std::string& n);
};

}  // namespace zg
