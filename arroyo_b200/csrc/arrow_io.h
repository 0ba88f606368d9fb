// Minimal Arrow C Data Interface import / export for 64-bit fixed-width columns and one level of
// struct nesting (the `window{start,end}` column, arroyo-planner/src/schemas.rs:7-23).
// This is the same mechanism the reference uses for its UDF dylib boundary
// (arroyo-udf/arroyo-udf-common/src/lib.rs:12-69).
#pragma once

#include <string>
#include <vector>

#include "common.cuh"

namespace ab {

// ---- import -------------------------------------------------------------------------------
struct InColumn {
  const uint64_t* data;  // first logical element (offsets applied)
  std::string format;
};

inline bool format_is_64bit(const char* f) {
  if (!f) return false;
  if (!strcmp(f, "l") || !strcmp(f, "L") || !strcmp(f, "g")) return true;
  if (!strncmp(f, "tsn:", 4)) return true;  // timestamp[ns]
  if (!strcmp(f, "tDn")) return true;       // duration[ns]
  return false;
}

// Validates a record batch exported as a struct array and returns its columns.
inline std::vector<InColumn> import_batch(const ArrowArray* a, const ArrowSchema* s, int64_t* n_rows) {
  AB_REQUIRE(a && s, ARROYO_B200_INVALID_ARGUMENT, "null batch or schema");
  AB_REQUIRE(s->format && !strcmp(s->format, "+s"), ARROYO_B200_INVALID_ARGUMENT,
             "batch must be exported as a struct array (format +s)");
  AB_REQUIRE(a->n_children == s->n_children, ARROYO_B200_INVALID_ARGUMENT, "array/schema children mismatch");
  AB_REQUIRE(a->null_count <= 0 || a->n_buffers == 0 || a->buffers[0] == nullptr, ARROYO_B200_UNSUPPORTED,
             "null rows at the struct level are not supported");
  std::vector<InColumn> cols;
  cols.reserve(a->n_children);
  for (int64_t i = 0; i < a->n_children; ++i) {
    const ArrowArray* c = a->children[i];
    const ArrowSchema* cs = s->children[i];
    AB_REQUIRE(c && cs, ARROYO_B200_INVALID_ARGUMENT, "null child");
    if (!format_is_64bit(cs->format)) {
      throw Error(ARROYO_B200_UNSUPPORTED, std::string("column ") + (cs->name ? cs->name : "?") +
                                               ": unsupported type '" + (cs->format ? cs->format : "") +
                                               "' (supported: l, L, g, tsn:)");
    }
    AB_REQUIRE(c->length >= a->length + a->offset, ARROYO_B200_INVALID_ARGUMENT, "child shorter than batch");
    AB_REQUIRE(c->n_buffers == 2, ARROYO_B200_INVALID_ARGUMENT, "primitive column must have 2 buffers");
    if (c->null_count != 0 && c->buffers[0] != nullptr) {
      // null_count may be -1 (unknown): count the zero validity bits of the logical range.
      int64_t nulls = c->null_count;
      if (nulls < 0) {
        const uint8_t* v = (const uint8_t*)c->buffers[0];
        nulls = 0;
        for (int64_t r = 0; r < a->length; ++r) {
          int64_t bit = c->offset + a->offset + r;
          if (!((v[bit >> 3] >> (bit & 7)) & 1)) ++nulls;
        }
      }
      if (nulls > 0)
        throw Error(ARROYO_B200_UNSUPPORTED, std::string("column ") + (cs->name ? cs->name : "?") +
                                                 " has nulls: NULL keys/values are outside the supported subset");
    }
    AB_REQUIRE(a->length == 0 || c->buffers[1] != nullptr, ARROYO_B200_INVALID_ARGUMENT, "null data buffer");
    InColumn ic;
    ic.data = (const uint64_t*)c->buffers[1] + c->offset + a->offset;
    ic.format = cs->format;
    cols.push_back(ic);
  }
  *n_rows = a->length;
  return cols;
}

// ---- export -------------------------------------------------------------------------------
struct OutColumn {
  std::string name;
  std::string format;               // "l", "g", "tsn:", or "+s" for a struct of children
  void* data = nullptr;             // pinned buffer from PinnedPool (ownership moves to the array)
  std::vector<OutColumn> children;  // struct only
  bool nullable = false;
  void* validity = nullptr;         // optional pinned validity bitmap
  int64_t null_count = 0;
};

namespace detail {
struct ArrayPriv {
  std::vector<const void*> buffers;
  std::vector<ArrowArray> child_storage;
  std::vector<ArrowArray*> child_ptrs;
  std::vector<void*> owned;  // pinned buffers to give back
};
struct SchemaPriv {
  std::string format, name;
  std::vector<ArrowSchema> child_storage;
  std::vector<ArrowSchema*> child_ptrs;
};

inline void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  auto* p = (ArrayPriv*)a->private_data;
  for (auto& c : p->child_storage)
    if (c.release) c.release(&c);
  for (void* b : p->owned) PinnedPool::get().free(b);
  delete p;
  a->release = nullptr;
}
inline void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  auto* p = (SchemaPriv*)s->private_data;
  for (auto& c : p->child_storage)
    if (c.release) c.release(&c);
  delete p;
  s->release = nullptr;
}

inline void fill_array(const OutColumn& col, int64_t n_rows, ArrowArray* out) {
  auto* p = new ArrayPriv();
  memset(out, 0, sizeof *out);
  out->length = n_rows;
  out->null_count = col.null_count;
  out->offset = 0;
  if (col.format == "+s") {
    p->buffers = {nullptr};
    p->child_storage.resize(col.children.size());
    for (size_t i = 0; i < col.children.size(); ++i) {
      fill_array(col.children[i], n_rows, &p->child_storage[i]);
    }
    for (auto& c : p->child_storage) p->child_ptrs.push_back(&c);
  } else {
    p->buffers = {col.validity, col.data};
    if (col.data) p->owned.push_back(col.data);
    if (col.validity) p->owned.push_back(col.validity);
  }
  out->n_buffers = (int64_t)p->buffers.size();
  out->buffers = p->buffers.data();
  out->n_children = (int64_t)p->child_ptrs.size();
  out->children = p->child_ptrs.empty() ? nullptr : p->child_ptrs.data();
  out->dictionary = nullptr;
  out->release = release_array;
  out->private_data = p;
}

inline void fill_schema(const OutColumn& col, ArrowSchema* out) {
  auto* p = new SchemaPriv();
  memset(out, 0, sizeof *out);
  p->format = col.format;
  p->name = col.name;
  p->child_storage.resize(col.children.size());
  for (size_t i = 0; i < col.children.size(); ++i) fill_schema(col.children[i], &p->child_storage[i]);
  for (auto& c : p->child_storage) p->child_ptrs.push_back(&c);
  out->format = p->format.c_str();
  out->name = p->name.c_str();
  out->metadata = nullptr;
  out->flags = col.nullable ? ARROW_FLAG_NULLABLE : 0;
  out->n_children = (int64_t)p->child_ptrs.size();
  out->children = p->child_ptrs.empty() ? nullptr : p->child_ptrs.data();
  out->dictionary = nullptr;
  out->release = release_schema;
  out->private_data = p;
}
}  // namespace detail

// Exports columns as one record batch (struct array + struct schema).
inline void export_batch(const std::vector<OutColumn>& cols, int64_t n_rows, ArrowArray* arr, ArrowSchema* sch) {
  OutColumn root;
  root.name = "";
  root.format = "+s";
  root.children = cols;
  detail::fill_array(root, n_rows, arr);
  detail::fill_schema(root, sch);
}

// Storage behind an ArroyoB200Batches value.
struct BatchesPriv {
  std::vector<ArrowArray> arrays;
  std::vector<ArrowSchema> schemas;
};

inline void batches_finish(BatchesPriv* p, ArroyoB200Batches* out) {
  out->n_batches = (int64_t)p->arrays.size();
  out->arrays = p->arrays.empty() ? nullptr : p->arrays.data();
  out->schemas = p->schemas.empty() ? nullptr : p->schemas.data();
  out->private_data = p;
}

inline void batches_release(ArroyoB200Batches* b) {
  if (!b || !b->private_data) return;
  auto* p = (BatchesPriv*)b->private_data;
  for (auto& a : p->arrays)
    if (a.release) a.release(&a);
  for (auto& s : p->schemas)
    if (s.release) s.release(&s);
  delete p;
  b->n_batches = 0;
  b->arrays = nullptr;
  b->schemas = nullptr;
  b->private_data = nullptr;
}

}  // namespace ab
