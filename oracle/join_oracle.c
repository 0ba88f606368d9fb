/* C restatement of the reference's instant (windowed) join -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Follows arroyo-worker/src/arrow/instant_join.rs:
 *   process_side (:109-172)   rows are routed to one join execution per distinct _timestamp; a batch older than the
 *                             operator's watermark is a panic (:129-139)
 *   handle_watermark (:256-283) every instant < watermark is finished in ascending order: a DataFusion HashJoinExec
 *                             over the rows of that instant (equi-join on one Int64 column; inner / left / right / full),
 *                             output = [left payload..., right payload..., _timestamp = max(l.ts, r.ts)]
 *                             (arroyo-planner/src/plan/join.rs:121-198)
 * HashJoinExec itself (datafusion-physical-plan 48.0.1, not under /root/reference) is restated from its published
 * algorithm: build a hash table on the left input, probe with the right input, emit matched pairs, then the unmatched
 * rows the join type keeps.  Row order is unspecified (tests compare multisets).
 *
 * All columns are int64 (timestamps in ns).  Payload = every column except _timestamp. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { JOIN_INNER = 0, JOIN_LEFT = 1, JOIN_RIGHT = 2, JOIN_FULL = 3 };

typedef struct {
  int n_cols, key_col, ts_col;
  int64_t n, cap;
  int64_t** cols; /* [n_cols][cap] */
} Side;

typedef struct JoinOracle {
  int join_type;
  Side side[2];
} JoinOracle;

typedef struct {
  int64_t n, cap;
  int n_cols;        /* left payload + right payload + 1 */
  int64_t** cols;    /* [n_cols][cap] */
  uint8_t** valid;   /* [n_cols][cap]: 0 = NULL (outer joins) */
} JoinOut;

static uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

JoinOracle* oracle_join_create(int join_type, int n_left_cols, int left_key_col, int left_ts_col, int n_right_cols,
                               int right_key_col, int right_ts_col) {
  JoinOracle* j = (JoinOracle*)calloc(1, sizeof *j);
  j->join_type = join_type;
  const int nc[2] = {n_left_cols, n_right_cols}, kc[2] = {left_key_col, right_key_col}, tc[2] = {left_ts_col, right_ts_col};
  for (int s = 0; s < 2; ++s) {
    j->side[s].n_cols = nc[s];
    j->side[s].key_col = kc[s];
    j->side[s].ts_col = tc[s];
    j->side[s].cols = (int64_t**)calloc((size_t)nc[s], sizeof(int64_t*));
  }
  return j;
}

void oracle_join_destroy(JoinOracle* j) {
  if (!j) return;
  for (int s = 0; s < 2; ++s) {
    for (int c = 0; c < j->side[s].n_cols; ++c) free(j->side[s].cols[c]);
    free(j->side[s].cols);
  }
  free(j);
}

/* instant_join.rs:109-172.  Returns -1 for the reference's panic (a row older than the watermark), else 0. */
int oracle_join_process(JoinOracle* j, int s, const int64_t* const* cols, int64_t n, int has_wm, int64_t wm) {
  Side* sd = &j->side[s];
  if (has_wm)
    for (int64_t i = 0; i < n; ++i)
      if (cols[sd->ts_col][i] < wm) return -1;
  if (sd->n + n > sd->cap) {
    int64_t cap = sd->cap ? sd->cap : 1024;
    while (cap < sd->n + n) cap *= 2;
    for (int c = 0; c < sd->n_cols; ++c) sd->cols[c] = (int64_t*)realloc(sd->cols[c], (size_t)cap * sizeof(int64_t));
    sd->cap = cap;
  }
  for (int c = 0; c < sd->n_cols; ++c) memcpy(sd->cols[c] + sd->n, cols[c], (size_t)n * sizeof(int64_t));
  sd->n += n;
  return 0;
}

JoinOut* oracle_join_out_create(const JoinOracle* j) {
  JoinOut* o = (JoinOut*)calloc(1, sizeof *o);
  o->n_cols = (j->side[0].n_cols - 1) + (j->side[1].n_cols - 1) + 1;
  o->cols = (int64_t**)calloc((size_t)o->n_cols, sizeof(int64_t*));
  o->valid = (uint8_t**)calloc((size_t)o->n_cols, sizeof(uint8_t*));
  return o;
}
void oracle_join_out_clear(JoinOut* o) { o->n = 0; }
void oracle_join_out_destroy(JoinOut* o) {
  if (!o) return;
  for (int c = 0; c < o->n_cols; ++c) {
    free(o->cols[c]);
    free(o->valid[c]);
  }
  free(o->cols);
  free(o->valid);
  free(o);
}

static void out_reserve(JoinOut* o, int64_t extra) {
  if (o->n + extra <= o->cap) return;
  int64_t cap = o->cap ? o->cap : 1024;
  while (cap < o->n + extra) cap *= 2;
  for (int c = 0; c < o->n_cols; ++c) {
    o->cols[c] = (int64_t*)realloc(o->cols[c], (size_t)cap * sizeof(int64_t));
    o->valid[c] = (uint8_t*)realloc(o->valid[c], (size_t)cap);
  }
  o->cap = cap;
}

/* one output row from left row li (or -1) and right row ri (or -1) */
static void emit_pair(const JoinOracle* j, JoinOut* o, int64_t li, int64_t ri) {
  out_reserve(o, 1);
  const int64_t r = o->n++;
  int oc = 0;
  int64_t lts = INT64_MIN, rts = INT64_MIN;
  for (int s = 0; s < 2; ++s) {
    const Side* sd = &j->side[s];
    const int64_t idx = s == 0 ? li : ri;
    for (int c = 0; c < sd->n_cols; ++c) {
      if (c == sd->ts_col) {
        if (idx >= 0) {
          if (s == 0) lts = sd->cols[c][idx];
          else rts = sd->cols[c][idx];
        }
        continue;
      }
      o->cols[oc][r] = idx >= 0 ? sd->cols[c][idx] : 0;
      o->valid[oc][r] = idx >= 0;
      ++oc;
    }
  }
  o->cols[oc][r] = lts > rts ? lts : rts;
  o->valid[oc][r] = 1;
}

typedef struct { int64_t ts, idx; } TsIdx;
static int cmp_tsidx(const void* a, const void* b) {
  const TsIdx *x = (const TsIdx*)a, *y = (const TsIdx*)b;
  if (x->ts != y->ts) return x->ts < y->ts ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* eligible rows (ts < wm) of a side sorted by (ts, arrival) */
static TsIdx* eligible(const Side* sd, int64_t wm, int64_t* n_out) {
  TsIdx* v = (TsIdx*)malloc((size_t)(sd->n > 0 ? sd->n : 1) * sizeof(TsIdx));
  int64_t n = 0;
  for (int64_t i = 0; i < sd->n; ++i)
    if (sd->cols[sd->ts_col][i] < wm) {
      v[n].ts = sd->cols[sd->ts_col][i];
      v[n].idx = i;
      ++n;
    }
  qsort(v, (size_t)n, sizeof(TsIdx), cmp_tsidx);
  *n_out = n;
  return v;
}

/* HashJoinExec over one instant: left rows l[0..nl), right rows r[0..nr) (indices into the sides) */
static void join_instant(const JoinOracle* j, const TsIdx* l, int64_t nl, const TsIdx* r, int64_t nr, JoinOut* o) {
  const Side *L = &j->side[0], *R = &j->side[1];
  const int keep_l = j->join_type == JOIN_LEFT || j->join_type == JOIN_FULL;
  const int keep_r = j->join_type == JOIN_RIGHT || j->join_type == JOIN_FULL;
  /* build on the left: chained hash table of left rows */
  uint64_t cap = 16;
  while (cap < (uint64_t)nl * 2 + 1) cap <<= 1;
  int64_t* head = (int64_t*)malloc((size_t)cap * sizeof(int64_t));
  int64_t* next = (int64_t*)malloc((size_t)(nl > 0 ? nl : 1) * sizeof(int64_t));
  uint8_t* l_matched = (uint8_t*)calloc((size_t)(nl > 0 ? nl : 1), 1);
  for (uint64_t i = 0; i < cap; ++i) head[i] = -1;
  for (int64_t i = nl - 1; i >= 0; --i) { /* reverse insertion keeps chains in arrival order */
    const uint64_t h = mix64((uint64_t)L->cols[L->key_col][l[i].idx]) & (cap - 1);
    next[i] = head[h];
    head[h] = i;
  }
  /* probe with the right */
  for (int64_t q = 0; q < nr; ++q) {
    const int64_t key = R->cols[R->key_col][r[q].idx];
    int any = 0;
    for (int64_t i = head[mix64((uint64_t)key) & (cap - 1)]; i >= 0; i = next[i])
      if (L->cols[L->key_col][l[i].idx] == key) {
        emit_pair(j, o, l[i].idx, r[q].idx);
        l_matched[i] = 1;
        any = 1;
      }
    if (!any && keep_r) emit_pair(j, o, -1, r[q].idx);
  }
  if (keep_l)
    for (int64_t i = 0; i < nl; ++i)
      if (!l_matched[i]) emit_pair(j, o, l[i].idx, -1);
  free(head);
  free(next);
  free(l_matched);
}

static void compact(Side* sd, int64_t wm) {
  int64_t w = 0;
  for (int64_t i = 0; i < sd->n; ++i)
    if (sd->cols[sd->ts_col][i] >= wm) {
      if (w != i)
        for (int c = 0; c < sd->n_cols; ++c) sd->cols[c][w] = sd->cols[c][i];
      ++w;
    }
  sd->n = w;
}

/* instant_join.rs:256-283: finish every instant < wm, oldest first */
void oracle_join_handle_watermark(JoinOracle* j, int64_t wm, JoinOut* out) {
  int64_t nl, nr;
  TsIdx* l = eligible(&j->side[0], wm, &nl);
  TsIdx* r = eligible(&j->side[1], wm, &nr);
  int64_t a = 0, b = 0;
  while (a < nl || b < nr) {
    int64_t t;
    if (a < nl && b < nr) t = l[a].ts < r[b].ts ? l[a].ts : r[b].ts;
    else t = a < nl ? l[a].ts : r[b].ts;
    int64_t a1 = a, b1 = b;
    while (a1 < nl && l[a1].ts == t) ++a1;
    while (b1 < nr && r[b1].ts == t) ++b1;
    join_instant(j, l + a, a1 - a, r + b, b1 - b, out);
    a = a1;
    b = b1;
  }
  free(l);
  free(r);
  compact(&j->side[0], wm);
  compact(&j->side[1], wm);
}

int64_t oracle_join_buffered(const JoinOracle* j, int side) { return j->side[side].n; }
