/* Compiled-C drive of the libzgpu C ABI (include/zgpu.h): what a cgo caller does, with no ctypes in
 * between. Built with plain gcc by tests/test_cabi_harness.py and run on a GPU box.
 *   part 1: the INTEGRATION.md snippet, verbatim semantics (write -> check -> HAS)
 *   part 2: 32 threads issue zg_check_bulk_str concurrently (pkg/authz/check.go:77-93 runs one goroutine
 *           per rule check); every answer must match the single-threaded answer, a failing call's message
 *           must be the caller's own (zg_last_error is thread-local), and the batcher must have coalesced
 *   part 3: error paths: unknown definition in a write, NULL arguments, ZG_E2BIG sizing protocol
 * Exit code 0 = all good; any other = the failing line is printed. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "zgpu.h"

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "FAIL %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, zg_last_error()); \
      exit(1);                                                             \
    }                                                                      \
  } while (0)

static const char *kSchema =
    "definition user {}\n"
    "definition namespace {\n"
    "  relation creator: user\n"
    "  relation viewer: user\n"
    "  permission view = viewer + creator\n"
    "  permission no_one_at_all = nil\n"
    "}\n";

enum { kThreads = 32, kPerThread = 64, kRounds = 40, kNamespaces = 200, kUsers = 16 };

typedef struct {
  zg_engine *e;
  int id;
  int failures;
} worker_arg;

static void name_of(char *buf, const char *prefix, int i) { sprintf(buf, "%s%d", prefix, i); }

/* expected: user u may view namespace n iff (n + u) % 3 == 0 (creator) or (n * 7 + u) % 5 == 0 (viewer) */
static int expected(int n, int u) { return ((n + u) % 3 == 0) || ((n * 7 + u) % 5 == 0); }

static void *worker(void *p) {
  worker_arg *a = (worker_arg *)p;
  char ns[kPerThread][24], us[kPerThread][24];
  zg_rel_str items[kPerThread];
  uint8_t out[kPerThread];
  int want[kPerThread];
  unsigned seed = 1234u + (unsigned)a->id * 77u;
  for (int r = 0; r < kRounds; ++r) {
    for (int i = 0; i < kPerThread; ++i) {
      seed = seed * 1664525u + 1013904223u;
      const int n = (int)((seed >> 8) % kNamespaces), u = (int)((seed >> 20) % kUsers);
      name_of(ns[i], "ns", n);
      name_of(us[i], "u", u);
      items[i].res_type = "namespace";
      items[i].res_id = ns[i];
      items[i].relation = (i % 17 == 16) ? "no_one_at_all" : "view";
      items[i].subj_type = "user";
      items[i].subj_id = us[i];
      items[i].subj_rel = "";
      want[i] = (i % 17 == 16) ? 0 : expected(n, u);
    }
    if (zg_check_bulk_str(a->e, items, kPerThread, out) != ZG_OK) {
      fprintf(stderr, "thread %d: %s\n", a->id, zg_last_error());
      a->failures++;
      continue;
    }
    for (int i = 0; i < kPerThread; ++i)
      if (out[i] != (want[i] ? ZG_HAS_PERMISSION : ZG_NO_PERMISSION)) a->failures++;
    /* a failing call on THIS thread: the message must be this thread's, whatever the others are doing */
    if (r % 8 == a->id % 8) {
      char tag[32];
      sprintf(tag, "nosuchtype%d", a->id);
      zg_update_str bad = {{tag, "x", "viewer", "user", "y", ""}, 0, ZG_OP_TOUCH};
      char msg[256];
      if (zg_write_relationships(a->e, &bad, 1, NULL, 0) != ZG_EINVAL) a->failures++;
      zg_last_error_copy(msg, sizeof msg);
      if (!strstr(msg, tag) || !strstr(zg_last_error(), tag)) {
        fprintf(stderr, "thread %d got someone else's error text: %s\n", a->id, msg);
        a->failures++;
      }
    }
  }
  return NULL;
}

int main(void) {
  /* ---- part 1: INTEGRATION.md */
  zg_engine *e;
  zg_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.device = 0;
  CHECK(zg_engine_create(&cfg, &e) == ZG_OK);
  CHECK(zg_load_schema(e, kSchema, strlen(kSchema)) == ZG_OK);
  zg_update_str up = {{"namespace", "ns1", "creator", "user", "paul", ""}, 0, ZG_OP_TOUCH};
  CHECK(zg_write_relationships(e, &up, 1, NULL, 0) == ZG_OK);
  zg_rel_str q = {"namespace", "ns1", "view", "user", "paul", ""};
  uint8_t code = 0;
  CHECK(zg_check_bulk_str(e, &q, 1, &code) == ZG_OK);
  CHECK(code == ZG_HAS_PERMISSION);
  zg_rel_str q2 = {"namespace", "ns1", "view", "user", "chani", ""};
  CHECK(zg_check_bulk_str(e, &q2, 1, &code) == ZG_OK && code == ZG_NO_PERMISSION);
  zg_rel_str q3 = {"namespace", "ns1", "nosuchperm", "user", "paul", ""};
  CHECK(zg_check_bulk_str(e, &q3, 1, &code) == ZG_OK && code == ZG_ITEM_ERROR);

  /* ---- part 2: concurrent callers */
  {
    char ns[24], us[24];
    for (int n = 0; n < kNamespaces; ++n)
      for (int u = 0; u < kUsers; ++u) {
        name_of(ns, "ns", n);
        name_of(us, "u", u);
        if ((n + u) % 3 == 0) {
          zg_update_str w = {{"namespace", ns, "creator", "user", us, ""}, 0, ZG_OP_TOUCH};
          CHECK(zg_write_relationships(e, &w, 1, NULL, 0) == ZG_OK);
        }
        if ((n * 7 + u) % 5 == 0) {
          zg_update_str w = {{"namespace", ns, "viewer", "user", us, ""}, 0, ZG_OP_TOUCH};
          CHECK(zg_write_relationships(e, &w, 1, NULL, 0) == ZG_OK);
        }
      }
  }
  pthread_t th[kThreads];
  worker_arg args[kThreads];
  for (int t = 0; t < kThreads; ++t) {
    args[t].e = e;
    args[t].id = t;
    args[t].failures = 0;
    CHECK(pthread_create(&th[t], NULL, worker, &args[t]) == 0);
  }
  int failures = 0;
  for (int t = 0; t < kThreads; ++t) {
    pthread_join(th[t], NULL);
    failures += args[t].failures;
  }
  CHECK(failures == 0);
  zg_stats st;
  CHECK(zg_stats_get(e, &st) == ZG_OK);
  printf("checks %llu launches %llu coalesced_launches %llu coalesced_requests %llu\n", (unsigned long long)st.checks,
         (unsigned long long)st.launches, (unsigned long long)st.coalesced_launches,
         (unsigned long long)st.coalesced_requests);
  CHECK(st.checks >= (unsigned long long)kThreads * kRounds * kPerThread);

  /* ---- part 3: error paths and the sizing protocol */
  CHECK(zg_check_bulk_str(e, NULL, 3, &code) == ZG_EINVAL);
  CHECK(zg_check_bulk(e, NULL, 0, NULL) == ZG_OK);
  size_t need = 0;
  uint64_t n_out = 0;
  int rc = zg_lookup_resources_str(e, "namespace", "view", "user", "u0", "", NULL, 0, &need, &n_out);
  CHECK(rc == ZG_E2BIG && need > 1 && n_out > 0);
  char *buf = (char *)malloc(need);
  CHECK(zg_lookup_resources_str(e, "namespace", "view", "user", "u0", "", buf, need, &need, &n_out) == ZG_OK);
  {
    uint64_t lines = 0, want = 0;
    for (char *c = buf; *c; ++c) lines += *c == '\n';
    for (int n = 0; n < kNamespaces; ++n) want += expected(n, 0);
    CHECK(lines == n_out && n_out == want);
  }
  free(buf);
  CHECK(zg_clear_relationships(e) == ZG_OK);
  CHECK(zg_check_bulk_str(e, &q, 1, &code) == ZG_OK && code == ZG_NO_PERMISSION);
  zg_engine_destroy(e);
  printf("cabi harness ok\n");
  return 0;
}
