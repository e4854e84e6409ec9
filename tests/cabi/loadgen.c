/* loadgen.c -- a load generator for the C ABI: N client threads (what goroutines are in the proxy: one per rule check,
 * pkg/authz/check.go:77-93, one per list request, pkg/authz/responsefilterer.go:165) against ONE zg_engine. Used by
 * scripts/cfg5_replay.py through ctypes so that the clients are native threads, not Python threads holding the GIL
 * between calls. Test / measurement infrastructure: not part of libzgpu.so. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "zgpu.h"

typedef struct {
  zg_engine *e;
  const zg_check *items; /* this client's list (post-filter shape): n_items checks */
  uint64_t n_items;
  uint8_t *out;          /* n_items answers */
  int rounds;            /* bulk calls per client */
  int do_lookup;         /* one LookupResources (pre-filter shape) */
  uint16_t res_type, perm, stype;
  uint32_t subj;
  uint32_t *ids;         /* lookup answer buffer */
  uint64_t ids_cap;
  uint64_t n_found;      /* out: ids the lookup returned */
  int rc;                /* out: first failing return code */
} lg_client;

static pthread_barrier_t g_start;
static int g_mode; /* 0 checks, 1 lookups, 2 lookup then checks */

static void *client_main(void *p) {
  lg_client *c = (lg_client *)p;
  pthread_barrier_wait(&g_start);
  if (g_mode != 0 && c->do_lookup) {
    /* a buffer large enough for the answer in one call (the retry protocol's answer cache holds 128 lookups, fewer
     * than the clients of this run) */
    uint64_t n = 0;
    int rc = zg_lookup_resources(c->e, c->res_type, c->perm, c->stype, c->subj, ZG_SREL_NONE, c->ids, c->ids_cap, &n);
    if (rc == ZG_E2BIG) {
      uint32_t *ids = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
      rc = zg_lookup_resources(c->e, c->res_type, c->perm, c->stype, c->subj, ZG_SREL_NONE, ids, n, &n);
      free(ids);
    }
    if (rc && !c->rc) c->rc = rc;
    c->n_found = n;
  }
  if (g_mode != 1)
    for (int r = 0; r < c->rounds; ++r) {
      int rc = zg_check_bulk(c->e, c->items, c->n_items, c->out);
      if (rc && !c->rc) c->rc = rc;
    }
  return NULL;
}

/* Runs n clients concurrently (all released by one barrier); returns the wall time in seconds, < 0 on failure. */
double loadgen_run(lg_client *clients, int n, int mode) {
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 256 * 1024);
  g_mode = mode;
  pthread_barrier_init(&g_start, NULL, (unsigned)n + 1);
  for (int i = 0; i < n; ++i)
    if (pthread_create(&th[i], &attr, client_main, &clients[i])) return -1.0;
  struct timespec t0, t1;
  pthread_barrier_wait(&g_start);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < n; ++i) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_barrier_destroy(&g_start);
  free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- filtered kube lists: every client hands one List body to zg_list_postfilter (mode 0: scan, resolve, ONE bulk
 * check, splice; pkg/authz/postfilter.go:17-178) or zg_list_prefilter (mode 1: LookupResources + scan + keep + splice;
 * pkg/authz/lookups.go:44-132). Bodies may be shared between clients (they are read-only); outputs are per client. */
typedef struct {
  zg_engine *e;
  const char *body;
  uint64_t body_len;
  zg_list_template tpl;
  char *out;
  uint64_t out_cap;
  uint64_t out_len; /* out: bytes of the last filtered body */
  int rounds;
  int rc;
} lg_list_client;

static void *list_client_main(void *p) {
  lg_list_client *c = (lg_list_client *)p;
  pthread_barrier_wait(&g_start);
  for (int r = 0; r < c->rounds; ++r) {
    size_t n = 0;
    int rc = g_mode == 0 ? zg_list_postfilter(c->e, c->body, c->body_len, &c->tpl, 1, c->out, c->out_cap, &n)
                         : zg_list_prefilter(c->e, c->body, c->body_len, ZG_LIST_ITEMS, &c->tpl, c->out, c->out_cap, &n);
    if (rc && !c->rc) c->rc = rc;
    c->out_len = n;
  }
  return NULL;
}

double loadgen_run_lists(lg_list_client *clients, int n, int mode) {
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 1024 * 1024);
  g_mode = mode;
  pthread_barrier_init(&g_start, NULL, (unsigned)n + 1);
  for (int i = 0; i < n; ++i)
    if (pthread_create(&th[i], &attr, list_client_main, &clients[i])) return -1.0;
  struct timespec t0, t1;
  pthread_barrier_wait(&g_start);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < n; ++i) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_barrier_destroy(&g_start);
  free(th);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
