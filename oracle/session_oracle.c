/* C restatement of the reference's session-window aggregate -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Follows arroyo-worker/src/arrow/session_aggregating_window.rs statement by statement:
 *   process_batch (:850-895)          drop rows older than the watermark (:858-868), lexsort by (key, _timestamp), split
 *                                     into per-key runs, KeyComputingHolder::add_batch for each
 *   ActiveSession::add_batch (:425-492)        including the scan's off-by-one: the row that ends the scan is still sent
 *                                     to the session (:464-479), and the first row of a run never extends data_end there
 *   KeyComputingHolder::{fill_active_session (:610-643), watermark_update (:557-603), add_batch (:645-677)}
 *   results_at_watermark (:99-160)    every key whose next action lies before the watermark is advanced
 *   output (:316-382)                 [key, window {start, end = data_end + gap}, aggregates, _timestamp = end - 1]
 * The session's Single-mode DataFusion aggregate is COUNT(*) / SUM / AVG / MIN / MAX over one Int64 column here.
 * One optional Int64 key column; rows are (key, value, ts). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int64_t start; int64_t off; int32_t len; int32_t next; } Node; /* one run under its start time */
typedef struct {
  int64_t key;
  int active;
  int64_t data_start, data_end;
  uint64_t rows;
  int64_t sum, mn, mx;
  int32_t head; /* pending runs ordered by (start, insertion) */
} KeyState;

typedef struct {
  int64_t *key, *start, *end, *rows, *sum, *mn, *mx, *ts;
  double* avg;
  int64_t n, cap;
} SessionOut;

typedef struct SessionOracle {
  int64_t gap;
  int keyed;
  int has_wm;
  int64_t wm;
  /* key -> state */
  int64_t* tab; /* state index + 1, 0 = empty */
  uint64_t tab_cap;
  KeyState* st;
  int64_t n_st, cap_st;
  Node* nodes;
  int64_t n_nodes, cap_nodes;
  int64_t *r_ts, *r_val;
  int64_t n_rows, cap_rows;
  uint64_t late_rows;
  int error; /* 1: flushed while adding (:672-675), 2: batch before data_start - gap (:447-451) */
} SessionOracle;

static uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

SessionOracle* oracle_session_create(int64_t gap, int keyed) {
  SessionOracle* s = (SessionOracle*)calloc(1, sizeof *s);
  s->gap = gap;
  s->keyed = keyed;
  s->tab_cap = 1024;
  s->tab = (int64_t*)calloc(s->tab_cap, sizeof(int64_t));
  return s;
}
void oracle_session_destroy(SessionOracle* s) {
  if (!s) return;
  free(s->tab);
  free(s->st);
  free(s->nodes);
  free(s->r_ts);
  free(s->r_val);
  free(s);
}
int oracle_session_error(const SessionOracle* s) { return s->error; }
uint64_t oracle_session_late_rows(const SessionOracle* s) { return s->late_rows; }

static void tab_grow(SessionOracle* s) {
  const uint64_t cap = s->tab_cap * 2;
  int64_t* t = (int64_t*)calloc(cap, sizeof(int64_t));
  for (int64_t i = 0; i < s->n_st; ++i) {
    uint64_t p = mix64((uint64_t)s->st[i].key) & (cap - 1);
    while (t[p]) p = (p + 1) & (cap - 1);
    t[p] = i + 1;
  }
  free(s->tab);
  s->tab = t;
  s->tab_cap = cap;
}

static KeyState* state_of(SessionOracle* s, int64_t key) {
  if ((uint64_t)(s->n_st + 1) * 2 > s->tab_cap) tab_grow(s);
  uint64_t p = mix64((uint64_t)key) & (s->tab_cap - 1);
  while (s->tab[p]) {
    KeyState* k = &s->st[s->tab[p] - 1];
    if (k->key == key) return k;
    p = (p + 1) & (s->tab_cap - 1);
  }
  if (s->n_st == s->cap_st) {
    s->cap_st = s->cap_st ? s->cap_st * 2 : 1024;
    s->st = (KeyState*)realloc(s->st, (size_t)s->cap_st * sizeof(KeyState));
  }
  KeyState* k = &s->st[s->n_st];
  memset(k, 0, sizeof *k);
  k->key = key;
  k->head = -1;
  s->tab[p] = ++s->n_st;
  return k;
}

static int32_t new_node(SessionOracle* s, int64_t start, int64_t off, int32_t len) {
  if (s->n_nodes == s->cap_nodes) {
    s->cap_nodes = s->cap_nodes ? s->cap_nodes * 2 : 1024;
    s->nodes = (Node*)realloc(s->nodes, (size_t)s->cap_nodes * sizeof(Node));
  }
  Node* n = &s->nodes[s->n_nodes];
  n->start = start;
  n->off = off;
  n->len = len;
  n->next = -1;
  return (int32_t)s->n_nodes++;
}

/* by_start.entry(start).or_default().push(run): ordered by start, after the runs with the same start */
static void pending_insert(SessionOracle* s, KeyState* k, int32_t node) {
  const int64_t start = s->nodes[node].start;
  int32_t prev = -1, cur = k->head;
  while (cur >= 0 && s->nodes[cur].start <= start) {
    prev = cur;
    cur = s->nodes[cur].next;
  }
  s->nodes[node].next = cur;
  if (prev < 0) k->head = node;
  else s->nodes[prev].next = node;
}

static void merge_rows(SessionOracle* s, KeyState* k, int64_t off, int32_t lo, int32_t hi) {
  for (int32_t i = lo; i < hi; ++i) {
    const int64_t v = s->r_val[off + i];
    k->sum = (int64_t)((uint64_t)k->sum + (uint64_t)v);
    if (v < k->mn) k->mn = v;
    if (v > k->mx) k->mx = v;
  }
  k->rows += (uint64_t)(hi - lo);
}

/* ActiveSession::add_batch (:425-492): returns the index of the first row left outside the session, or len */
static int32_t active_add_batch(SessionOracle* s, KeyState* k, int64_t off, int32_t n) {
  const int64_t* ts = s->r_ts + off;
  const int64_t start = ts[0], end = ts[n - 1];
  if (end < k->data_end + s->gap) {
    if (end > k->data_end) k->data_end = end;
    if (start < k->data_start) k->data_start = start;
    merge_rows(s, k, off, 0, n);
    return n;
  }
  if (k->data_end + s->gap < start) return 0;
  if (start < k->data_start - s->gap) {
    s->error |= 2;
    return n;
  }
  if (start < k->data_start) k->data_start = start;
  int32_t index = 1;
  while (index < n) {
    const int64_t value = ts[index];
    ++index; /* the reference increments before testing: the row that ends the scan stays in the session */
    if (value < k->data_end) continue;
    if (value < k->data_end + s->gap) {
      k->data_end = value;
      continue;
    }
    break;
  }
  merge_rows(s, k, off, 0, index);
  return index;
}

/* KeyComputingHolder::fill_active_session (:610-643) */
static void fill_active_session(SessionOracle* s, KeyState* k) {
  while (k->head >= 0) {
    const int64_t first = s->nodes[k->head].start;
    if (k->data_end + s->gap < first) break;
    /* pop_first(): every run stored under this start, in insertion order */
    int32_t h = k->head, tail = h;
    while (s->nodes[tail].next >= 0 && s->nodes[s->nodes[tail].next].start == first) tail = s->nodes[tail].next;
    k->head = s->nodes[tail].next;
    s->nodes[tail].next = -1;
    for (int32_t node = h; node >= 0;) {
      const int32_t next = s->nodes[node].next;
      const int64_t off = s->nodes[node].off;
      const int32_t len = s->nodes[node].len;
      const int32_t rem = active_add_batch(s, k, off, len);
      if (rem < len) pending_insert(s, k, new_node(s, s->r_ts[off + rem], off + rem, len - rem));
      node = next;
    }
  }
}

static void out_push(SessionOut* o, const SessionOracle* s, const KeyState* k) {
  if (o->n == o->cap) {
    o->cap = o->cap ? o->cap * 2 : 1024;
#define GROW(f, T) o->f = (T*)realloc(o->f, (size_t)o->cap * sizeof(T))
    GROW(key, int64_t); GROW(start, int64_t); GROW(end, int64_t); GROW(rows, int64_t); GROW(sum, int64_t);
    GROW(mn, int64_t); GROW(mx, int64_t); GROW(ts, int64_t); GROW(avg, double);
#undef GROW
  }
  const int64_t i = o->n++;
  const int64_t end = k->data_end + s->gap;
  o->key[i] = k->key;
  o->start[i] = k->data_start;
  o->end[i] = end;
  o->rows[i] = (int64_t)k->rows;
  o->sum[i] = k->sum;
  o->mn[i] = k->mn;
  o->mx[i] = k->mx;
  o->avg[i] = (double)k->sum / (double)k->rows;
  o->ts[i] = end - 1;
}

/* KeyComputingHolder::watermark_update (:557-603); returns the number of sessions finished */
static int watermark_update(SessionOracle* s, KeyState* k, int64_t wm, SessionOut* out) {
  int flushed = 0;
  for (;;) {
    if (k->active) {
      if (k->data_end + s->gap < wm) {
        if (out) out_push(out, s, k);
        k->active = 0;
        ++flushed;
      } else {
        break;
      }
    } else {
      if (k->head < 0) break;
      const int64_t initial = s->nodes[k->head].start;
      if ((__int128)wm + s->gap < (__int128)initial) break;
      k->active = 1;
      k->data_start = k->data_end = initial;
      k->rows = 0;
      k->sum = 0;
      k->mn = INT64_MAX;
      k->mx = INT64_MIN;
      fill_active_session(s, k);
    }
  }
  return flushed;
}

typedef struct { int64_t key, ts, val; } Row;
static int cmp_row(const void* a, const void* b) {
  const Row *x = (const Row*)a, *y = (const Row*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  if (x->ts != y->ts) return x->ts < y->ts ? -1 : 1;
  return 0;
}
/* stable sort by (key, ts): merge sort via qsort on (key, ts, original index) */
typedef struct { Row r; int64_t idx; } RowI;
static int cmp_rowi(const void* a, const void* b) {
  const RowI *x = (const RowI*)a, *y = (const RowI*)b;
  const int c = cmp_row(&x->r, &y->r);
  if (c) return c;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* process_batch (:850-895) + add_at_watermark (:162-231) */
void oracle_session_process_batch(SessionOracle* s, const int64_t* key, const int64_t* val, const int64_t* ts, int64_t n,
                                  int has_wm, int64_t wm) {
  s->has_wm = has_wm; /* ctx.last_present_watermark() at the time of the call */
  s->wm = wm;
  RowI* rows = (RowI*)malloc((size_t)(n > 0 ? n : 1) * sizeof(RowI));
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (s->has_wm && ts[i] < s->wm) {
      ++s->late_rows;
      continue;
    }
    rows[m].r.key = s->keyed ? key[i] : 0;
    rows[m].r.ts = ts[i];
    rows[m].r.val = val ? val[i] : 0;
    rows[m].idx = i;
    ++m;
  }
  qsort(rows, (size_t)m, sizeof(RowI), cmp_rowi);
  if (s->n_rows + m > s->cap_rows) {
    int64_t cap = s->cap_rows ? s->cap_rows : 4096;
    while (cap < s->n_rows + m) cap *= 2;
    s->r_ts = (int64_t*)realloc(s->r_ts, (size_t)cap * sizeof(int64_t));
    s->r_val = (int64_t*)realloc(s->r_val, (size_t)cap * sizeof(int64_t));
    s->cap_rows = cap;
  }
  for (int64_t a = 0; a < m;) {
    int64_t b = a + 1;
    while (b < m && rows[b].r.key == rows[a].r.key) ++b;
    const int64_t off = s->n_rows;
    for (int64_t i = a; i < b; ++i) {
      s->r_ts[s->n_rows] = rows[i].r.ts;
      s->r_val[s->n_rows] = rows[i].r.val;
      ++s->n_rows;
    }
    KeyState* k = state_of(s, rows[a].r.key);
    /* KeyComputingHolder::add_batch (:645-677) */
    pending_insert(s, k, new_node(s, s->r_ts[off], off, (int32_t)(b - a)));
    if (s->has_wm) {
      if (k->active) fill_active_session(s, k);
      if (watermark_update(s, k, s->wm, NULL)) s->error |= 1;
    }
    a = b;
  }
  free(rows);
}

SessionOut* oracle_session_out_create(void) { return (SessionOut*)calloc(1, sizeof(SessionOut)); }
void oracle_session_out_clear(SessionOut* o) { o->n = 0; }
void oracle_session_out_destroy(SessionOut* o) {
  if (!o) return;
  free(o->key); free(o->start); free(o->end); free(o->rows); free(o->sum); free(o->mn); free(o->mx); free(o->ts); free(o->avg);
  free(o);
}

/* results_at_watermark (:99-160): keys whose next watermark action is before the watermark */
void oracle_session_handle_watermark(SessionOracle* s, int64_t wm, SessionOut* out) {
  for (int64_t i = 0; i < s->n_st; ++i) {
    KeyState* k = &s->st[i];
    if (!k->active && k->head < 0) continue;
    const __int128 action = k->active ? (__int128)k->data_end + s->gap : (__int128)s->nodes[k->head].start - s->gap;
    if (!(action < (__int128)wm)) continue;
    watermark_update(s, k, wm, out);
  }
}
